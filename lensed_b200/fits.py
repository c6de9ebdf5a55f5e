"""Minimal uncompressed-FITS image reader / writer (numpy only).

Lensed reads its image, weight, mask and PSF through CFITSIO
(reference src/data.c:36-127) and writes the result layers as a multi-extension
FITS file (src/data.c:129-165, 372-391).  CFITSIO is not part of this build; the
hot path only needs plain 2-D image HDUs, so this module implements exactly
that: primary/IMAGE HDUs with BITPIX 8/16/32/64/-32/-64, BSCALE/BZERO, the
``file.fits[x0:x1,y0:y1]`` section syntax used by
``examples/test_sersic_bulge.ini:2`` and the ``[extname]``/``[n]`` HDU selector.
"""
from __future__ import annotations

import re
import numpy as np

_BLOCK = 2880
_DTYPES = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


def _parse_header(buf: bytes, pos: int):
    """Parse one header unit starting at ``pos``; returns (cards, new_pos)."""
    cards = {}
    while True:
        block = buf[pos:pos + _BLOCK]
        if len(block) < _BLOCK:
            raise ValueError("truncated FITS header")
        pos += _BLOCK
        done = False
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:]
            if val.lstrip().startswith("'"):
                m = re.match(r"\s*'((?:[^']|'')*)'", val)
                cards[key] = m.group(1).replace("''", "'").rstrip() if m else ""
            else:
                val = val.split("/")[0].strip()
                if val in ("T", "F"):
                    cards[key] = val == "T"
                else:
                    try:
                        cards[key] = int(val)
                    except ValueError:
                        try:
                            cards[key] = float(val.replace("D", "E"))
                        except ValueError:
                            cards[key] = val
        if done:
            return cards, pos


def read_hdus(path: str):
    """Return a list of (header dict, 2-D float64/native array or None)."""
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0
    hdus = []
    while pos < len(buf):
        try:
            hdr, pos = _parse_header(buf, pos)
        except ValueError:
            break
        naxis = hdr.get("NAXIS", 0)
        shape = [hdr[f"NAXIS{i}"] for i in range(naxis, 0, -1)]
        bitpix = hdr.get("BITPIX", 8)
        count = int(np.prod(shape)) if naxis else 0
        nbytes = count * abs(bitpix) // 8 + hdr.get("PCOUNT", 0)
        data = None
        if count:
            raw = np.frombuffer(buf, dtype=_DTYPES[bitpix], count=count, offset=pos)
            data = raw.reshape(shape)
            bscale, bzero = hdr.get("BSCALE", 1), hdr.get("BZERO", 0)
            if bscale != 1 or bzero != 0:
                data = data.astype(np.float64) * bscale + bzero
        pos += (nbytes + _BLOCK - 1) // _BLOCK * _BLOCK
        hdus.append((hdr, data))
    return hdus


_SPEC = re.compile(r"^(?P<file>[^\[\]]+)(?P<rest>(\[[^\]]*\])*)$")


def read_image(spec: str, dtype=np.float32):
    """Read a 2-D image given a CFITSIO-style file spec.

    Supports ``name.fits``, ``name.fits[ext]`` and the pixel section
    ``name.fits[x0:x1,y0:y1]`` (1-based, inclusive).  Returns
    ``(array[height, width], pcs)`` with ``pcs = (rx, ry, sx, sy)`` the pixel
    coordinate system of the section as reference src/data.c:167-234 derives
    it (origin of the cut-out, unit scale), so that model coordinates keep
    referring to the full frame.
    """
    m = _SPEC.match(spec)
    if not m:
        raise ValueError(f"bad FITS spec: {spec}")
    brackets = re.findall(r"\[([^\]]*)\]", m.group("rest") or "")
    hdus = read_hdus(m.group("file"))
    section = None
    ext = None
    for b in brackets:
        if re.match(r"^[\d\s:,*-]+$", b) and ("," in b):
            section = b
        else:
            ext = b
    if ext is None:
        hdr, data = next((h, d) for h, d in hdus if d is not None and d.ndim == 2)
    elif ext.strip().isdigit():
        hdr, data = hdus[int(ext)]
    else:
        hdr, data = next((h, d) for h, d in hdus if str(h.get("EXTNAME", "")).upper() == ext.strip().upper())
    if data is None or data.ndim != 2:
        raise ValueError(f"{spec}: no 2-D image")
    rx, ry, sx, sy = 1.0, 1.0, 1.0, 1.0
    if section:
        axes = section.split(",")
        if len(axes) != 2:
            raise ValueError(f"{spec}: image section must have two axes")
        (x0, x1, xi), (y0, y1, yi) = (_section_range(spec, a) for a in axes)
        data = data[_section_slice(y0, y1, yi), _section_slice(x0, x1, xi)]
        # src/data.c:262-270: origin = first pixel of the section, scale = +-increment
        rx, sx = float(x0), float(xi if x1 > x0 else -xi)
        ry, sy = float(y0), float(yi if y1 > y0 else -yi)
    return np.ascontiguousarray(data, dtype=dtype), (rx, ry, sx, sy)


def _section_range(spec: str, text: str):
    """`first:last[:increment]` of one axis of a CFITSIO image section
    (1-based, inclusive; first > last reverses the axis).  The `*` / `-*`
    shorthands are rejected: the reference derives its pixel origin from the
    section's first pixel (src/data.c:262-270), which CFITSIO leaves unset for them."""
    parts = [t.strip() for t in text.split(":")]
    if any("*" in t for t in parts):
        raise ValueError(f"{spec}: '*' in an image section is not supported, write the pixel range out (first:last)")
    try:
        vals = [int(t) for t in parts]
    except ValueError:
        raise ValueError(f"{spec}: bad image section '{text}'") from None
    if len(vals) == 2:
        vals.append(1)
    if len(vals) != 3 or vals[0] < 1 or vals[1] < 1 or vals[2] < 1:
        raise ValueError(f"{spec}: bad image section '{text}'")
    return tuple(vals)


def _section_slice(first: int, last: int, inc: int):
    if last >= first:
        return slice(first - 1, last, inc)
    stop = last - 2
    return slice(first - 1, stop if stop >= 0 else None, -inc)


def _card(key: str, value, comment: str = "") -> bytes:
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20}"
    elif isinstance(value, (float, np.floating)):
        v = f"{float(value):>20.12G}"
    else:
        v = f"'{str(value):<8}'"
        v = f"{v:<20}"
    s = f"{key:<8}= {v}"
    if comment:
        s += f" / {comment}"
    return f"{s:<80}"[:80].encode("ascii")


def _header(cards) -> bytes:
    hdr = b"".join(cards + [f"{'END':<80}".encode("ascii")])
    return hdr + b" " * (-len(hdr) % _BLOCK)


def write_layers(path: str, layers, names, empty_primary: bool = True):
    """Write float32 image layers in the structure of the reference's results
    file (write_fits, src/data.c:129-165): an empty primary HDU (16-bit, no
    axes) carrying ORIGIN and DATE, then one IMAGE extension per layer with its
    EXTNAME -- so IMG is HDU 1 and PVL HDU 6 of a results file, as for the
    reference.  empty_primary=False puts the first layer into the primary HDU
    instead (a plain image file)."""
    import datetime
    out = bytearray()
    if empty_primary:
        out += _header([_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", 16), _card("NAXIS", 0),
                        _card("EXTEND", True), _card("ORIGIN", "lensed-b200", "FITS file originator"),
                        _card("DATE", datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%S"),
                              "file creation date (YYYY-MM-DDThh:mm:ss UT)")])
    for idx, (img, name) in enumerate(zip(layers, names)):
        img = np.asarray(img, dtype=np.float32)
        h, w = img.shape
        primary = idx == 0 and not empty_primary
        cards = [_card("SIMPLE", True, "conforms to FITS standard") if primary else _card("XTENSION", "IMAGE", "IMAGE extension")]
        cards += [_card("BITPIX", -32), _card("NAXIS", 2), _card("NAXIS1", w), _card("NAXIS2", h)]
        cards += [_card("EXTEND", True)] if primary else [_card("PCOUNT", 0), _card("GCOUNT", 1)]
        cards.append(_card("EXTNAME", name, "extension name"))
        data = img.astype(">f4").tobytes()
        data += b"\0" * (-len(data) % _BLOCK)
        out += _header(cards) + data
    with open(path, "wb") as f:
        f.write(bytes(out))
