// homography_b200.mjs — Node.js face of the engine: the `Homography` class surface of Eric-Canas/Homography.js
// (constructor, setReferencePoints, setSourcePoints, setDestinyPoints, setImage, setTriangles, warp,
// getTransformationMatrixAsCSS, transformHTMLElement) with every
// arithmetic step executed on the GPU through the N-API addon (js/hgwarp_napi.c -> libhgwarp.so).
//
// This file is the JavaScript twin of ../homography.py (same state machine, same engine calls); the Python twin
// is the one the test-suite drives, because the build image has no JavaScript engine.  Reference line numbers
// (H.js:<n>) point at the behaviour being reproduced: normalised-range auto-detection (no value > 8.0), in-place
// (de)normalisation of caller-owned typed arrays, cache invalidation, the forward / inverse dispatch thresholds
// and the bare-string throws.
import { createRequire } from 'node:module';
const native = createRequire(import.meta.url)('./hgwarp.node');

const NORMALIZED_MAX = 8.0; // H.js:36
const MAX_CSS_DECIMAL = 5;  // H.js:31
const KIND = { affine: 0, projective: 1 };
const positive = (v) => v !== null && v > 0;
const isTyped = (a) => ArrayBuffer.isView(a);

function asPointArray(points) { // H.js:220 / 339
  return isTyped(points) ? points : new Float32Array(points.flat());
}
function anyAbove(arr, limit) { // H.js:1539
  for (let i = 0; i < arr.length; i++) if (arr[i] > limit) return true;
  return false;
}
function scaleInPlace(p, sx, sy, divide) { // H.js:1603 / 1621
  for (let i = 0; i < p.length; i++) {
    const s = (i & 1) === 0 ? sx : sy;
    p[i] = divide ? p[i] / s : p[i] * s;
  }
}
function minMaxXY(p) { // H.js:1558, unrounded
  let mnx = Infinity, mny = Infinity, mxx = -Infinity, mxy = -Infinity;
  for (let i = 0; i < p.length; i++) {
    const v = p[i];
    if ((i & 1) === 0) { if (v > mxx) mxx = v; if (v < mnx) mnx = v; }
    else { if (v > mxy) mxy = v; if (v < mny) mny = v; }
  }
  return [mnx, mny, mxx, mxy];
}
function selectTransform(first, points) { // H.js:1444
  const n = points.length;
  switch (first) {
    case 'auto':
      if (n === 6) return 'affine';
      if (n === 8) return 'projective';
      if (n > 8) return 'piecewiseaffine';
      throw (`Transforms must contain at least 3 points but only ${n / 2} were given`);
    case 'piecewiseaffine':
      if (n < 6) throw (`A piecewise (or affine) transform needs to determine least three reference points but only ${n / 2} were given`);
      return first;
    case 'affine':
      if (n !== 6) throw (`An affine transform needs to determine exactly three reference points but ${n / 2} were given`);
      return first;
    case 'projective':
      if (n !== 8) throw (`A projective transform needs to determine exactly four reference points but ${n / 2} were given`);
      return first;
    default:
      throw (`Transform "${first}" is unknown`);
  }
}
const f64 = (a) => (a instanceof Float64Array ? a : Float64Array.from(a));
const f32 = (a) => (a instanceof Float32Array ? a : Float32Array.from(a));

class Homography {
  constructor(transform = 'auto', width = null, height = null, { device = 0, triangulate = null } = {}) {
    this._ctx = native.createContext(device);
    // `triangulate(points) -> Uint32Array` (optional): the reference calls delaunator@5.0.0 here (H.js:1216).  Without
    // it the addon's own restatement of that package runs (hg_delaunay); pass `(p) => new Delaunator(p).triangles`
    // to use the real package, or call setTriangles().
    this._triangulate = triangulate;
    this._width = width === null ? null : Math.round(width);
    this._height = height === null ? null : Math.round(height);
    this._objectiveWidth = this._objectiveHeight = null;
    this._xOutputOffset = this._yOutputOffset = null;
    this._srcPoints = this._dstPoints = null;
    this.firstTransformSelected = this.transform = transform.toLowerCase();
    this._image = null;
    this._minSrcX = this._minSrcY = this._maxSrcX = this._maxSrcY = null;
    this._srcPointsAreNormalized = this._dstPointsAreNormalized = true;
    this._mapState = null;        // what the reference's shared _trianglesCorrespondencesMatrix would hold
    this._triangles = this._initialTriangles = null;
    this._transformMatrix = this._piecewiseMatrices = null;
    this._meshOnDevice = false;
  }

  setReferencePoints(srcPoints, dstPoints, image = null, width = null, height = null, srcNorm = null, dstNorm = null) {
    if (typeof srcPoints === 'undefined' || typeof dstPoints === 'undefined')
      throw ('Source and Destiny points must be defined when calling setReferencePoints().');
    this._dstPoints = null;
    this.setSourcePoints(srcPoints, image, width, height, srcNorm);
    this.setDestinyPoints(dstPoints, dstNorm);
  }

  setSourcePoints(points, image = null, width = null, height = null, pointsAreNormalized = null) {
    const pts = asPointArray(points);
    this._srcPoints = pts;
    this._meshOnDevice = false;
    this._srcPointsAreNormalized = pointsAreNormalized === null ? !anyAbove(pts, NORMALIZED_MAX) : pointsAreNormalized;
    this._transformMatrix = null;
    this.transform = selectTransform(this.firstTransformSelected, pts);
    this._objectiveWidth = this._objectiveHeight = null;
    if (image !== null) this.setImage(image, width, height);
    else if (width !== null || height !== null) this._setSrcWidthHeight(width, height);
    if (this._width !== null && this._height !== null && this._srcPointsAreNormalized) this._denormalizeSrc();
    if (this._dstPoints !== null && this.transform !== 'piecewiseaffine')
      this._transformMatrix = this._solve(this._srcPoints, this._dstPoints).matrix;
    if (this.transform === 'piecewiseaffine' && this._mapState === null) {
      this._triangles = this._initialTriangles;
      this._piecewiseMatrices = null;
      if (!this._srcPointsAreNormalized || (positive(this._width) && positive(this._height))) this._setPiecewiseParameters();
      else if (this._triangles === null) this._triangles = this._delaunay(this._srcPoints);
    }
  }

  setImage(image, width = null, height = null) { // ImageData-like {data, width, height} (H.js:292-299)
    if (!image || !isTyped(image.data)) throw ('setImage() needs an ImageData-like object ({data, width, height})');
    this._image = image.data;
    native.setImage(this._ctx, image.data, image.width, image.height); // stays device-resident
    this._setSrcWidthHeight(image.width, image.height);
    if (this._srcPoints !== null && this.transform === 'piecewiseaffine') this._setPiecewiseParameters();
    if (this._dstPoints !== null && (this._objectiveWidth <= 0 || this._objectiveHeight <= 0)) this._induceObjective();
  }

  setDestinyPoints(points, pointsAreNormalized = null) {
    const pts = asPointArray(points);
    if (this._srcPoints !== null && pts.length !== this._srcPoints.length)
      throw (`It must be the same amount of destiny points (${pts.length / 2}) than source points (${this._srcPoints.length / 2})`);
    this._dstPoints = pts;
    this._dstPointsAreNormalized = pointsAreNormalized === null ? !anyAbove(pts, NORMALIZED_MAX) : pointsAreNormalized;
    const haveSize = positive(this._width) && positive(this._height);
    let limitsDone = false;
    if (this.transform !== 'piecewiseaffine') {
      if (this._dstPointsAreNormalized && haveSize && this.transform === 'projective') this._denormalizeDst();
      this._sameRange();
      const r = this._solve(this._srcPoints, this._dstPoints, this._image !== null);
      this._transformMatrix = r.matrix;
      if (this._image !== null) { this._setLimits(r.limits); limitsDone = true; }
    } else {
      this._piecewiseMatrices = null;
    }
    if (!limitsDone && (this._image !== null || (this.transform === 'piecewiseaffine' && haveSize))) this._induceObjective();
    if (this.transform === 'piecewiseaffine' && haveSize) {
      if (this._dstPointsAreNormalized) this._denormalizeDst();
      this._setPiecewiseParameters();
    }
  }

  setTriangles(triangles) { // H.js:517
    this._triangles = triangles;
    this._meshOnDevice = false;
    if ((!this._srcPointsAreNormalized || (positive(this._width) && positive(this._height))) && this._srcPoints !== null)
      this._setPiecewiseParameters();
  }

  warp(image = null, asHTMLPromise = false, applyAlwaysInverse = false) { // H.js:408
    if (asHTMLPromise) throw ('asHTMLPromise needs a DOM; only ImageData results exist outside a browser');
    if (image !== null) this.setImage(image);
    else if (this._image === null)
      throw ('warp() must receive an image if it was not setted before through `setImage(img)` or  `setSourcePoints(points, img)`');
    const oW = this._objectiveWidth, oH = this._objectiveHeight, W = this._width, H = this._height;
    const empty = !(oW * oH >= 1) || Number.isNaN(oW * oH); // H.js:436-441
    let data;
    if (this.transform === 'piecewiseaffine') {
      const inverse = applyAlwaysInverse || (oW > W || oH > H || oW * 1.2 < W || oH * 1.2 < H); // H.js:421
      data = inverse ? this._inversePiecewise(empty) : this._forwardPiecewise(empty);
    } else if (this.transform === 'affine') {
      const inverse = applyAlwaysInverse || (oW !== W || oH !== H); // H.js:426
      data = inverse ? this._inverseGeometric(empty) : this._forwardGeometric(empty);
    } else {
      data = this._inverseGeometric(empty);
    }
    if (empty) return { data: new Uint8ClampedArray(4), width: 1, height: 1 };
    return { data, width: oW, height: oH };
  }

  // String form of the current affine / projective matrix for the CSS `transform` property (H.js:548-586): pure
  // string work on the matrix the engine solved (six float32 values -> `matrix(...)`, eight doubles laid out column-major
  // with the z row / column of the identity -> `matrix3d(...)`), every value through toFixed(5) (maxCSSDecimal, H.js:31).
  getTransformationMatrixAsCSS(srcPoints = null, dstPoints = null, width = null, height = null) {
    if (width !== null || height !== null) this._setSrcWidthHeight(width, height);
    if (srcPoints !== null) this.setSourcePoints(srcPoints, null, width, height);
    if (dstPoints !== null) this.setDestinyPoints(dstPoints);
    if (this._srcPoints === null) throw ('Impossible to calculate a transform when srcPoints are not set');
    else if (this._dstPoints === null) throw ('Impossible to calculate a transform when dstPoints are not set');
    else if (this._transformMatrix === null) throw ('Transform matrix can not be calculated');
    const m = this._transformMatrix, D = MAX_CSS_DECIMAL;
    if (this.transform === 'affine') return `matrix(${Array.from(m, (v) => v.toFixed(D)).join(', ')})`;
    if (this.transform === 'projective') {
      const cells = [];
      let i = 0;
      for (let dy = 0; dy < 4; dy++) {
        for (let dx = 0; dx < 4; dx++) {
          if ((dy === 2 && dx === 2) || (dy === 3 && dx === 3)) cells.push('1');
          else if (dy === 2 || dx === 2) cells.push('0');
          else cells.push(m[((i++) * 3) % 8].toFixed(D)); // h0 h3 h6 | h1 h4 h7 | h2 h5: the 3x3 transposed
        }
      }
      return `matrix3d(${cells.join(', ')})`;
    }
    throw (`Only "affine" or "projective" transforms can be applied on the CSS transform property, but ${this.transform} selected`);
  }

  // H.js:611 — any object with getBoundingClientRect() and a style (a DOM element in a browser / jsdom)
  transformHTMLElement(element, srcPoints = null, dstPoints = null) {
    const rect = element.getBoundingClientRect();
    element.style.transform = this.getTransformationMatrixAsCSS(srcPoints, dstPoints, rect.width, rect.height);
  }

  // ---------------------------------------------------------------- state plumbing (H.js:637-896)
  _solve(src, dst, withLimits = true) {
    const kind = KIND[this.transform];
    if (kind === undefined) throw (`${this.transform} transform does not exist`);
    return native.solveWithLimits(this._ctx, kind, f64(src), f64(dst), withLimits ? this._width : 1, withLimits ? this._height : 1);
  }
  _setLimits(l) { [this._xOutputOffset, this._yOutputOffset, this._objectiveWidth, this._objectiveHeight] = l; }
  _denormalizeSrc() { scaleInPlace(this._srcPoints, this._width, this._height, false); this._srcPointsAreNormalized = false; this._meshOnDevice = false; }
  _denormalizeDst() { scaleInPlace(this._dstPoints, this._width, this._height, false); this._dstPointsAreNormalized = false; }
  _delaunay(points) {
    // H.js:1216: new Delaunator(points).triangles.  Default: the library's restatement of delaunator 5.0.0
    // (hg_delaunay); {triangulate} overrides it, e.g. with the real package.
    return this._triangulate ? this._triangulate(points) : native.delaunay(points);
  }

  _setSrcWidthHeight(width, height) {
    const changed = this._width !== width || this._height !== height;
    this._width = width; this._height = height;
    if (!changed) return;
    this._width = Math.round(width); this._height = Math.round(height);
    this._mapState = null;
    if (this.transform === 'projective') {
      if (this._srcPoints !== null && this._srcPointsAreNormalized) this._denormalizeSrc();
      if (this._dstPoints !== null && this._dstPointsAreNormalized) this._denormalizeDst();
      if (this._dstPoints !== null && this._srcPoints !== null) {
        const r = this._solve(this._srcPoints, this._dstPoints);
        this._transformMatrix = r.matrix; this._setLimits(r.limits);
      }
    }
    if (this._srcPoints !== null && this.transform === 'piecewiseaffine') this._setPiecewiseParameters();
  }

  _induceObjective() { // H.js:693
    if (this.transform === 'affine' || this.transform === 'projective') {
      if (this._transformMatrix === null && this._srcPointsAreNormalized !== this._dstPointsAreNormalized) this._sameRange();
      const r = this._solve(this._srcPoints, this._dstPoints);
      if (this._transformMatrix === null) this._transformMatrix = r.matrix;
      this._setLimits(r.limits);
    } else if (!this._dstPointsAreNormalized) {
      const [a, b, c, d] = minMaxXY(this._dstPoints);
      this._xOutputOffset = Math.round(a); this._yOutputOffset = Math.round(b);
      this._objectiveWidth = Math.round(c) - this._xOutputOffset;   // difference of ROUNDED extrema (H.js:707-710)
      this._objectiveHeight = Math.round(d) - this._yOutputOffset;
    } else if (positive(this._width) && positive(this._height)) {
      const [a, b, c, d] = minMaxXY(this._dstPoints);
      this._xOutputOffset = Math.round(a); this._yOutputOffset = Math.round(b);
      this._objectiveWidth = Math.round((c - a) * this._width);
      this._objectiveHeight = Math.round((d - b) * this._height);
    } else {
      throw ('Trying to calculate a the output width and height of a Piecewise Affine transform but source width and height are not set');
    }
  }

  _setPiecewiseParameters() { // H.js:738
    if (this._srcPoints === null) throw ('Trying to set the Piecewise Affine Transform parameters before setting the Source Points.');
    if (this._triangles === null) { this._triangles = this._delaunay(this._srcPoints); this._meshOnDevice = false; }
    if (this._srcPointsAreNormalized) {
      if (positive(this._width) && positive(this._height)) this._denormalizeSrc();
      else throw ('Trying to set the Piecewise Affine Transform parameters without knowing the source points ranges');
    }
    if (this._mapState === null) {
      const [a, b, c, d] = minMaxXY(this._srcPoints);
      this._minSrcX = Math.round(a); this._minSrcY = Math.round(b); this._maxSrcX = Math.round(c); this._maxSrcY = Math.round(d);
      this._mapState = 'forward'; // the engine rasterises the forward map lazily, when a forward warp is dispatched
    }
    if (this._dstPoints !== null && this._piecewiseMatrices === null && this._triangles !== null) {
      if (this._dstPointsAreNormalized) this._denormalizeDst();
      if (this._srcPointsAreNormalized !== this._dstPointsAreNormalized) this._sameRange();
      this._uploadMesh();
      this._piecewiseMatrices = native.piecewiseMatrices(this._ctx, f32(this._dstPoints), this._triangles.length / 3);
    }
  }
  _uploadMesh() {
    if (!this._meshOnDevice) { native.setMesh(this._ctx, f32(this._srcPoints), Uint32Array.from(this._triangles)); this._meshOnDevice = true; }
  }

  _sameRange() { // H.js:876
    if (this._dstPointsAreNormalized === this._srcPointsAreNormalized) return;
    const haveSize = positive(this._width) && positive(this._height);
    if (this._dstPointsAreNormalized && haveSize) {
      scaleInPlace(this._srcPoints, this._width, this._height, true); this._srcPointsAreNormalized = true; this._meshOnDevice = false;
    } else if (this._srcPointsAreNormalized && haveSize) {
      this._denormalizeSrc();
    } else {
      throw ('Impossible to put source and destiny points in the same range. Possible solutions: \n' +
        '1. Give a source width/height when calling setSrcPoints.\n2. Set the input image before.\n' +
        '3. Give Source and Destiny points in the same range (both normalized or both in image dimensions)');
    }
  }

  // ---------------------------------------------------------------- the four loops -> device
  _inverseGeometric(empty) { // H.js:987
    this._sameRange();
    if (empty) return null;
    return native.warpInversePoints(this._ctx, KIND[this.transform], f64(this._dstPoints), f64(this._srcPoints),
      this._xOutputOffset, this._yOutputOffset, this._objectiveWidth, this._objectiveHeight);
  }
  _forwardGeometric(empty) { // H.js:911
    if (empty) return null;
    return native.warpForwardMatrix(this._ctx, KIND[this.transform], this._transformMatrix,
      this._xOutputOffset, this._yOutputOffset, this._objectiveWidth, this._objectiveHeight);
  }
  _inversePiecewise(empty) { // H.js:1029
    this._mapState = 'inverse';
    if (empty) return null;
    this._uploadMesh();
    return native.warpPiecewiseInverse(this._ctx, f32(this._dstPoints), this._xOutputOffset, this._yOutputOffset,
      this._objectiveWidth, this._objectiveHeight, this._minSrcX, this._minSrcY);
  }
  _forwardPiecewise(empty) { // H.js:948
    if (empty) return null;
    this._uploadMesh();
    return native.warpPiecewiseForward(this._ctx, f32(this._dstPoints), this._xOutputOffset, this._yOutputOffset,
      this._objectiveWidth, this._objectiveHeight, this._minSrcX, this._minSrcY, this._maxSrcX, this._maxSrcY,
      this._mapState === 'inverse' ? 1 : 0);
  }
}

// File I/O for hosts without a canvas (the reference reads pixels through drawImage + getImageData, H.js:1071-1076, and
// returns PNG data URLs, H.js:480-483): PNG bytes <-> the {data, width, height} form setImage() / warp() use.
function imageDataFromPNG(bytes) { return native.pngDecode(bytes); }
function imageDataFromJPEG(bytes) { return native.jpegDecode(bytes); } // baseline and progressive JPEG; throws for CMYK / arithmetic-coded files
function pngFromImageData(image) { return native.pngEncode(image.data, image.width, image.height); }

export { Homography, imageDataFromPNG, imageDataFromJPEG, pngFromImageData };
