"""Optional file hand-off between test and evaluation: the two volumes the reference's WriteHook
stores per subject — `{subject}_probabilities.nii.gz` (float32 foreground p) and `{subject}_prediction.nii.gz` (uint8
argmax), bin-dl/brats_test_default.py:90-108 — written without SimpleITK, in a background thread like the reference's
`thread.do_work(..., in_background=True)` (common/utils/threadhelper.py:9-11), so that bin-eval / bin-analysis keep
working on files while the metrics themselves come from the in-memory route (hooks.DeviceMetricsHook).

NIfTI-1 single-file layout (348-byte header + 4-byte extension flag, data at offset 352, gzip by file name).  Geometry
follows the ITK convention the reference's `conversion.ImageProperties` carry (origin / spacing / direction in LPS):
NIfTI stores RAS, i.e. the first two world axes negated, as qform (quaternion) and sform (affine rows), both with code 1.
numpy arrays are (z, y, x) like `sitk.GetArrayFromImage`; the file's fastest axis is x.

Not pinned against the reference's writer (SimpleITK is not installable here): the header fields are checked against
the NIfTI-1 specification and the files round-trip through `read_nifti` (tests/test_nifti.py).
"""
import gzip
import math
import os
import struct
import threading

import numpy as np

_DTYPES = {np.dtype(np.uint8): (2, 8), np.dtype(np.int16): (4, 16), np.dtype(np.int32): (8, 32), np.dtype(np.float32): (16, 32),
           np.dtype(np.float64): (64, 64), np.dtype(np.int8): (256, 8), np.dtype(np.uint16): (512, 16), np.dtype(np.uint32): (768, 32)}
_CODES = {code: dt for dt, (code, _) in _DTYPES.items()}


def _quaternion(rot):
    """(b, c, d) of the unit quaternion of a proper rotation matrix (nifti1_io's mat44_to_quatern, a >= 0)."""
    r = np.asarray(rot, dtype=np.float64)
    a = r[0, 0] + r[1, 1] + r[2, 2] + 1.0
    if a > 0.5:
        a = 0.5 * math.sqrt(a)
        b = 0.25 * (r[2, 1] - r[1, 2]) / a
        c = 0.25 * (r[0, 2] - r[2, 0]) / a
        d = 0.25 * (r[1, 0] - r[0, 1]) / a
    else:
        xd, yd, zd = 1.0 + r[0, 0] - (r[1, 1] + r[2, 2]), 1.0 + r[1, 1] - (r[0, 0] + r[2, 2]), 1.0 + r[2, 2] - (r[0, 0] + r[1, 1])
        if xd > 1.0:
            b = 0.5 * math.sqrt(xd)
            c, d, a = 0.25 * (r[0, 1] + r[1, 0]) / b, 0.25 * (r[0, 2] + r[2, 0]) / b, 0.25 * (r[2, 1] - r[1, 2]) / b
        elif yd > 1.0:
            c = 0.5 * math.sqrt(yd)
            b, d, a = 0.25 * (r[0, 1] + r[1, 0]) / c, 0.25 * (r[1, 2] + r[2, 1]) / c, 0.25 * (r[0, 2] - r[2, 0]) / c
        else:
            d = 0.5 * math.sqrt(zd)
            b, c, a = 0.25 * (r[0, 2] + r[2, 0]) / d, 0.25 * (r[1, 2] + r[2, 1]) / d, 0.25 * (r[1, 0] - r[0, 1]) / d
        if a < 0.0:
            b, c, d = -b, -c, -d
    return b, c, d


def _rotation(b, c, d, qfac):
    a = math.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    r = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    r[:, 2] *= qfac
    return r


def header_bytes(shape_zyx, dtype, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=(1, 0, 0, 0, 1, 0, 0, 0, 1)):
    """The 352 bytes in front of the voxel data (little endian)."""
    dtype = np.dtype(dtype)
    if dtype not in _DTYPES:
        raise ValueError('dtype {} has no NIfTI-1 code'.format(dtype))
    if len(shape_zyx) not in (2, 3):
        raise ValueError('2-D or 3-D volumes only, got shape {}'.format(tuple(shape_zyx)))
    dims = [int(s) for s in reversed(shape_zyx)] + [1] * (3 - len(shape_zyx))       # (nx, ny, nz)
    spacing = [float(s) for s in spacing] + [1.0] * (3 - len(spacing))
    origin = [float(o) for o in origin] + [0.0] * (3 - len(origin))
    d = np.asarray(direction, dtype=np.float64)
    d = d.reshape(2, 2) if d.size == 4 else d.reshape(3, 3)
    if d.shape == (2, 2):
        d3 = np.eye(3)
        d3[:2, :2] = d
        d = d3
    flip = np.diag([-1.0, -1.0, 1.0])                       # ITK LPS -> NIfTI RAS
    rot = flip @ d
    qfac = 1.0
    if np.linalg.det(rot) < 0:
        qfac = -1.0
        rot = rot.copy()
        rot[:, 2] *= -1.0
    qb, qc, qd = _quaternion(rot)
    affine = (flip @ d) * np.asarray(spacing)[None, :]
    offset = flip @ np.asarray(origin)
    code, bitpix = _DTYPES[dtype]
    h = bytearray(352)
    struct.pack_into('<i', h, 0, 348)
    struct.pack_into('<8h', h, 40, len(shape_zyx) if len(shape_zyx) == 3 else 2, dims[0], dims[1], dims[2] if len(shape_zyx) == 3 else 1, 1, 1, 1, 1)
    struct.pack_into('<hh', h, 70, code, bitpix)
    struct.pack_into('<8f', h, 76, qfac, spacing[0], spacing[1], spacing[2], 0.0, 0.0, 0.0, 0.0)
    struct.pack_into('<fff', h, 108, 352.0, 1.0, 0.0)        # vox_offset, scl_slope, scl_inter
    h[123] = 2                                                # xyzt_units: millimetres
    struct.pack_into('<hh', h, 252, 1, 1)                    # qform_code, sform_code: scanner anatomical
    struct.pack_into('<6f', h, 256, qb, qc, qd, offset[0], offset[1], offset[2])
    for row in range(3):
        struct.pack_into('<4f', h, 280 + 16 * row, affine[row, 0], affine[row, 1], affine[row, 2], offset[row])
    h[344:348] = b'n+1\x00'
    return bytes(h)


def write_nifti(path, array_zyx, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=(1, 0, 0, 0, 1, 0, 0, 0, 1), compresslevel=1):
    """Write a (z, y, x) [or (y, x)] numpy array as NIfTI-1; `.gz` file names are gzip-compressed."""
    a = np.ascontiguousarray(array_zyx)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    data = header_bytes(a.shape, a.dtype, spacing, origin, direction) + a.astype(a.dtype.newbyteorder('<'), copy=False).tobytes()
    tmp = path + '.part'
    if path.endswith('.gz'):
        with gzip.open(tmp, 'wb', compresslevel=compresslevel) as f:
            f.write(data)
    else:
        with open(tmp, 'wb') as f:
            f.write(data)
    os.replace(tmp, path)


def read_nifti(path):
    """(array (z, y, x), {'spacing', 'origin', 'direction'} in the ITK / LPS convention) of a NIfTI-1 single file."""
    opener = gzip.open if path.endswith('.gz') else open
    with opener(path, 'rb') as f:
        raw = f.read()
    endian = '<' if struct.unpack_from('<i', raw, 0)[0] == 348 else '>'
    if struct.unpack_from(endian + 'i', raw, 0)[0] != 348 or raw[344:347] != b'n+1':
        raise ValueError('{} is not a single-file NIfTI-1 image'.format(path))
    dim = struct.unpack_from(endian + '8h', raw, 40)
    code, _ = struct.unpack_from(endian + 'hh', raw, 70)
    pixdim = struct.unpack_from(endian + '8f', raw, 76)
    vox_offset, slope, inter = struct.unpack_from(endian + 'fff', raw, 108)
    qform_code, sform_code = struct.unpack_from(endian + 'hh', raw, 252)
    ndim = dim[0]
    shape = tuple(int(v) for v in reversed(dim[1:1 + ndim]))
    dt = _CODES[code].newbyteorder(endian)
    n = int(np.prod(shape))
    a = np.frombuffer(raw, dtype=dt, count=n, offset=int(vox_offset)).reshape(shape).astype(_CODES[code])
    if slope not in (0.0, 1.0) or inter != 0.0:
        a = a * slope + inter
    flip = np.diag([-1.0, -1.0, 1.0])
    if sform_code > 0:
        rows = np.array([struct.unpack_from(endian + '4f', raw, 280 + 16 * r) for r in range(3)], dtype=np.float64)
        affine, offset = rows[:, :3], rows[:, 3]
        spacing = np.linalg.norm(affine, axis=0)
        direction = flip @ (affine / spacing[None, :])
    else:
        qb, qc, qd, ox, oy, oz = struct.unpack_from(endian + '6f', raw, 256)
        direction = flip @ _rotation(qb, qc, qd, -1.0 if pixdim[0] < 0 else 1.0)
        spacing, offset = np.asarray(pixdim[1:4], dtype=np.float64), np.array([ox, oy, oz], dtype=np.float64)
    geometry = {'spacing': tuple(float(s) for s in spacing[:max(ndim, 2)]), 'origin': tuple(float(o) for o in (flip @ offset)[:max(ndim, 2)]),
                'direction': tuple(float(v) for v in direction.reshape(-1))}
    return a, geometry


class AsyncNiftiWriteHook:
    """TestLoopHook-protocol hook (common/trainloop/hooks.py:67-98; only on_test_subject_end / on_test_end act) that stores
    what the reference's WriteHook stores (bin-dl/brats_test_default.py:90-108), from numpy arrays or from the CUDA
    tensors assembly.DeviceSubjectAssembler hands over (one device->host copy per volume, then a background thread)."""

    def __init__(self, out_dir=None, probability_entry='probabilities', properties_entry='properties', subject_entry='subject') -> None:
        self.out_dir = out_dir
        self.probability_entry, self.properties_entry, self.subject_entry = probability_entry, properties_entry, subject_entry
        self.threads = []
        self.written = []

    def on_startup(self): pass
    def end_startup(self, context): pass
    def on_termination(self, context): pass
    def on_test_start(self, task_context, context): pass
    def on_test_batch_start(self, batch_context, task_context, context): pass
    def on_test_batch_end(self, batch_context, task_context, context): pass
    def on_test_subject_start(self, subject_context, task_context, context): pass

    def on_test_subject_end(self, subject_context, task_context, context):
        from .hooks import subject_outputs
        data = subject_context.subject_data
        p, prediction = subject_outputs(data[self.probability_entry], data.get('prediction'))
        if hasattr(p, 'is_cuda'):
            p, prediction = p.cpu().numpy(), prediction.cpu().numpy()
        props = data.get(self.properties_entry)
        geometry = {k: getattr(props, k) for k in ('spacing', 'origin', 'direction') if props is not None and hasattr(props, k)}
        subject = data.get(self.subject_entry, subject_context.subject_index)
        out_dir = self.out_dir if self.out_dir is not None else getattr(context, 'test_dir', '.')
        paths = (os.path.join(out_dir, '{}_probabilities.nii.gz'.format(subject)), os.path.join(out_dir, '{}_prediction.nii.gz'.format(subject)))

        def work():
            write_nifti(paths[0], p.astype(np.float32, copy=False), **geometry)
            write_nifti(paths[1], prediction.astype(np.uint8, copy=False), **geometry)
            self.written.extend(paths)
        t = threading.Thread(target=work, daemon=False)
        t.start()
        self.threads.append(t)

    def on_test_end(self, task_context, context):
        self.join()

    def join(self):
        for t in self.threads:
            t.join()
        self.threads = []
