"""Host-side image helpers on the INPUT side of the render pass: sRGB<->linear LUT
(color/srgb.go), texture load with gamma conversion (internal/imageutil/image.go:48-106) and
the fixed-point bilinear ``Resize`` used to build mip chains (internal/imageutil/resize.go).

These run once at scene-build time; the finished mip chain is an input of prc_scene_upload.
"""
from __future__ import annotations

import math as _m

import numpy as np

f32 = np.float32
LUT_SIZE = 1024


def _linear2srgb64(v: float) -> float:  # color/srgb.go:86-93
    return v * 12.92 if v <= 0.0031308 else 1.055 * _m.pow(v, 1.0 / 2.4) - 0.055


def _srgb2linear64(v: float) -> float:  # color/srgb.go:77-84
    return v / 12.92 if v <= 0.04045 else _m.pow((v + 0.055) / 1.055, 2.4)


# color/srgb.go:62-75
_LIN2SRGB = np.array([_linear2srgb64(i / LUT_SIZE) for i in range(LUT_SIZE)] + [0.0])
_LIN2SRGB[LUT_SIZE] = _LIN2SRGB[LUT_SIZE - 1]
_SRGB2LIN = np.array([_srgb2linear64(i / LUT_SIZE) for i in range(LUT_SIZE)] + [0.0])
_SRGB2LIN[LUT_SIZE] = _SRGB2LIN[LUT_SIZE - 1]


def _lut_lookup(lut: np.ndarray, v: np.ndarray, one_is_ge: bool) -> np.ndarray:
    """FromLinear2sRGB / FromsRGB2Linear with T=float32 (color/srgb.go:15-52)."""
    v = np.asarray(v, dtype=np.float32)
    i = (v * f32(LUT_SIZE)).astype(np.float32)
    ifloor = i.astype(np.int64) & (LUT_SIZE - 1)
    v0 = lut[ifloor].astype(np.float32)
    v1 = lut[ifloor + 1].astype(np.float32)
    fr = (i - ifloor.astype(np.float32)).astype(np.float32)
    out = (v0 * (f32(1.0) - fr) + v1 * fr).astype(np.float32)
    out = np.where(v <= 0, f32(0), out)
    out = np.where((v >= 1) if one_is_ge else (v == 1), f32(1), out)
    return out.astype(np.float32)


def linear_to_srgb(v):
    return _lut_lookup(_LIN2SRGB, v, one_is_ge=False)


def srgb_to_linear(v):
    return _lut_lookup(_SRGB2LIN, v, one_is_ge=True)


def gamma_lut_u8() -> np.ndarray:
    """u8 -> u8 table of shader.GammaCorrection (shader/gamma.go:13-18):
    uint8(FromLinear2sRGB(float32(c)/0xff)*0xff + 0.5)."""
    c = np.arange(256, dtype=np.float32) / f32(255)
    s = linear_to_srgb(c)
    return (s * f32(255) + f32(0.5)).astype(np.float32).astype(np.int64).astype(np.uint8)


def srgb_image_to_linear(pix: np.ndarray) -> np.ndarray:
    """imageutil.LoadImage with GammaCorrect(true) (image.go:76-103): RGB channels through
    FromsRGB2Linear with +0.5 rounding, alpha untouched."""
    table = (srgb_to_linear(np.arange(256, dtype=np.float32) / f32(255)) * f32(255) + f32(0.5)).astype(np.float32).astype(np.int64).astype(np.uint8)
    out = pix.copy()
    out[..., :3] = table[pix[..., :3]]
    return out


def load_image(path: str, gamma_correct: bool = False) -> np.ndarray:
    """Decode to RGBA8 [h, w, 4] (Go: image.Decode + draw.Draw into *image.RGBA)."""
    from PIL import Image

    im = Image.open(path)
    im = im.convert("RGBA")
    pix = np.asarray(im, dtype=np.uint8).copy()
    # image.RGBA is alpha-premultiplied; all reference fixtures on the path are opaque.
    if gamma_correct:
        pix = srgb_image_to_linear(pix)
    return pix


# ----------------------------------------------------------------------------- Resize
def _create_weights8(dy: int, filter_length: int, scale: np.float32):
    """createWeights8 with the `linear` kernel (resize.go:143-164)."""
    scale = f32(scale)
    filter_length = filter_length * int(max(_m.ceil(float(scale)), 1))
    filter_factor = f32(min(float(f32(1.0) / scale), 1.0))
    y = np.arange(dy, dtype=np.float32)
    interp = (scale * (y + f32(0.5)) - f32(0.5)).astype(np.float32)
    start = interp.astype(np.int64) - filter_length // 2 + 1
    interp = (interp - start.astype(np.float32)).astype(np.float32)
    i = np.arange(filter_length, dtype=np.float32)
    x = ((interp[:, None] - i[None, :]) * filter_factor).astype(np.float32)
    ax = np.abs(x)
    k = np.where(ax <= 1, (f32(1) - ax).astype(np.float32), f32(0)).astype(np.float32)
    coeffs = (k * f32(256)).astype(np.float32).astype(np.int64)  # int16(kernel(in)*256), truncation
    return coeffs, start, filter_length


def _resize_pass(src: np.ndarray, out_len: int, scale) -> np.ndarray:
    """resizeRGBA (resize.go:76-117): filters along axis 1 of src [rows, n, 4] and returns the
    TRANSPOSED result [out_len, rows, 4]."""
    rows, n, _ = src.shape
    coeffs, start, fl = _create_weights8(out_len, 2, scale)
    max_x = n - 1
    xi = start[:, None] + np.arange(fl)[None, :]
    xi = np.where((xi >= 0) & (xi < max_x), xi, np.where(xi >= max_x, max_x, 0))
    out = np.empty((out_len, rows, 4), dtype=np.uint8)
    s32 = src.astype(np.int64)
    for y in range(out_len):  # bounded by the new size
        c = coeffs[y]
        nz = c != 0
        acc = (s32[:, xi[y][nz], :] * c[nz][None, :, None]).sum(axis=1)
        tot = int(c[nz].sum())
        q = acc // tot  # Go int32 division truncates toward zero; operands are non-negative here
        out[y] = np.clip(q, 0, 255).astype(np.uint8)
    return out


def resize(width: int, height: int, img: np.ndarray) -> np.ndarray:
    """imageutil.Resize (resize.go:16-62): two-pass bilinear with int16*256 coefficients."""
    h0, w0 = img.shape[:2]
    sx = f32(w0) / f32(width)
    sy = f32(h0) / f32(height)
    if width == w0 and height == h0:
        return img
    temp = _resize_pass(img, width, sx)      # [width, h0, 4]  (transposed)
    result = _resize_pass(temp, height, sy)  # [height, width, 4]
    return result


def build_mipmap(img: np.ndarray) -> list[np.ndarray]:
    """buffer.NewTexture (buffer/texture.go:35-69): L = int(Log2(max(dx,dy)))+1 levels, each
    resized FROM LEVEL 0 to (dx / 2^i, dy / 2^i)."""
    dy, dx = img.shape[:2]
    if dx == 1 and dy == 1:
        return [img]
    L = int(f32(_m.log2(float(max(dx, dy))))) + 1
    mips = [img]
    for i in range(1, L):
        w = dx // int(2 ** i)
        h = dy // int(2 ** i)
        mips.append(resize(w, h, img))
    return mips
