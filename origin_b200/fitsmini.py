"""Minimal FITS image reader (numpy only).

ORIGIN reads its spectral dictionaries with ``astropy.io.fits`` (reference
``muse_origin/origin.py:515-533``: every HDU after the primary one is a 1-D
IMAGE extension holding one profile, with a ``FWHM`` header card).  astropy is
outside the hot path and is not a dependency of this package, so the few
hundred bytes of FITS needed to read those dictionaries (and label maps such as
``tests/segmap.fits``) are parsed here.

Only what the hot path needs: primary + IMAGE extensions, BITPIX in
{8, 16, 32, 64, -32, -64}, BSCALE/BZERO applied when present.
"""

import numpy as np

_BLOCK = 2880
_DTYPES = {8: 'u1', 16: '>i2', 32: '>i4', 64: '>i8', -32: '>f4', -64: '>f8'}


def _parse_value(raw):
    raw = raw.strip()
    if raw.startswith("'"):
        end = raw.find("'", 1)
        while end != -1 and raw[end:end + 2] == "''":
            end = raw.find("'", end + 2)
        return raw[1:end].rstrip()
    raw = raw.split('/')[0].strip()
    if raw in ('T', 'F'):
        return raw == 'T'
    try:
        return int(raw)
    except ValueError:
        try:
            return float(raw.replace('D', 'E'))
        except ValueError:
            return raw


def _read_header(buf, off):
    header = {}
    while True:
        block = buf[off:off + _BLOCK]
        if len(block) < _BLOCK:
            raise ValueError('truncated FITS header')
        off += _BLOCK
        done = False
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode('ascii', 'replace')
            key = card[:8].strip()
            if key == 'END':
                done = True
                break
            if card[8:10] == '= ':
                header[key] = _parse_value(card[10:])
        if done:
            return header, off


def read_hdus(path):
    """Return ``[(header_dict, ndarray_or_None), ...]`` for every HDU."""
    with open(path, 'rb') as f:
        buf = f.read()
    off, out = 0, []
    while off < len(buf):
        header, off = _read_header(buf, off)
        naxis = header.get('NAXIS', 0)
        shape = [header['NAXIS%d' % (i + 1)] for i in range(naxis)]
        bitpix = header.get('BITPIX', 8)
        count = int(np.prod(shape)) if naxis else 0
        nbytes = count * abs(bitpix) // 8
        nbytes += header.get('PCOUNT', 0)
        data = None
        if count:
            if header.get('XTENSION', 'IMAGE').strip() != 'IMAGE':
                raise ValueError('only IMAGE extensions are supported')
            arr = np.frombuffer(buf, dtype=_DTYPES[bitpix], count=count, offset=off)
            arr = arr.reshape(shape[::-1])
            arr = arr.astype(arr.dtype.newbyteorder('='))
            bscale, bzero = header.get('BSCALE', 1), header.get('BZERO', 0)
            if bscale != 1 or bzero != 0:
                arr = arr * bscale + bzero
            data = arr
        off += (nbytes + _BLOCK - 1) // _BLOCK * _BLOCK
        out.append((header, data))
    return out
