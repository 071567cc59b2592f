"""File hand-off: NIfTI-1 volumes without SimpleITK.  Header fields against the NIfTI-1
specification (nifti1.h offsets and codes), round trips, geometry conventions (ITK LPS <-> NIfTI RAS), the asynchronous
write hook with numpy and tensor subjects."""
import gzip
import struct
import types

import numpy as np
import pytest
import torch

from rcu_b200 import nifti


def test_header_fields_follow_the_nifti1_layout():
    h = nifti.header_bytes((155, 240, 200), np.float32, spacing=(1.0, 1.5, 2.0), origin=(10.0, -20.0, 30.0))
    assert len(h) == 352
    assert struct.unpack_from('<i', h, 0)[0] == 348                                     # sizeof_hdr
    assert struct.unpack_from('<8h', h, 40) == (3, 200, 240, 155, 1, 1, 1, 1)           # dim: x fastest
    assert struct.unpack_from('<hh', h, 70) == (16, 32)                                 # DT_FLOAT32, bitpix
    assert struct.unpack_from('<4f', h, 76) == (1.0, 1.0, 1.5, 2.0)                     # qfac, pixdim
    assert struct.unpack_from('<fff', h, 108) == (352.0, 1.0, 0.0)                      # vox_offset, scl_slope, scl_inter
    assert h[123] == 2 and struct.unpack_from('<hh', h, 252) == (1, 1)                  # mm; qform / sform codes
    # identity direction in LPS = 180 degrees about z in RAS: quaternion (a, b, c, d) = (0, 0, 0, 1)
    assert np.allclose(struct.unpack_from('<3f', h, 256), (0.0, 0.0, 1.0))
    assert struct.unpack_from('<3f', h, 268) == (-10.0, 20.0, 30.0)                     # qoffset: x, y negated
    assert struct.unpack_from('<4f', h, 280) == (-1.0, 0.0, 0.0, -10.0)                 # srow_x
    assert struct.unpack_from('<4f', h, 296) == (0.0, -1.5, 0.0, 20.0)                  # srow_y
    assert struct.unpack_from('<4f', h, 312) == (0.0, 0.0, 2.0, 30.0)                   # srow_z
    assert h[344:348] == b'n+1\x00' and h[348:352] == b'\x00\x00\x00\x00'               # magic, no extensions
    assert struct.unpack_from('<hh', nifti.header_bytes((4, 4), np.uint8), 70) == (2, 8)
    with pytest.raises(ValueError):
        nifti.header_bytes((4, 4), np.complex64)
    with pytest.raises(ValueError):
        nifti.header_bytes((2, 2, 2, 2), np.float32)


@pytest.mark.parametrize('dtype', [np.float32, np.uint8, np.int16, np.float64])
@pytest.mark.parametrize('gz', [True, False])
def test_round_trip_values_and_geometry(tmp_path, dtype, gz):
    rng = np.random.default_rng(3)
    a = (rng.random((7, 12, 9)) * 100).astype(dtype)
    th = 0.3
    rot = np.array([[np.cos(th), -np.sin(th), 0.0], [np.sin(th), np.cos(th), 0.0], [0.0, 0.0, 1.0]])
    path = str(tmp_path / ('vol.nii.gz' if gz else 'vol.nii'))
    nifti.write_nifti(path, a, spacing=(0.9, 1.1, 3.0), origin=(-12.5, 40.0, 7.25), direction=tuple(rot.reshape(-1)))
    b, geo = nifti.read_nifti(path)
    assert b.dtype == a.dtype and np.array_equal(a, b)
    assert np.allclose(geo['spacing'], (0.9, 1.1, 3.0), atol=1e-6) and np.allclose(geo['origin'], (-12.5, 40.0, 7.25), atol=1e-5)
    assert np.allclose(np.reshape(geo['direction'], (3, 3)), rot, atol=1e-6)
    if gz:
        with gzip.open(path, 'rb') as f:
            assert len(f.read()) == 352 + a.nbytes
    # the quaternion alone carries the same orientation (readers that ignore the sform)
    raw = gzip.open(path, 'rb').read() if gz else open(path, 'rb').read()
    qb, qc, qd = struct.unpack_from('<3f', raw, 256)
    flip = np.diag([-1.0, -1.0, 1.0])
    assert np.allclose(flip @ nifti._rotation(qb, qc, qd, 1.0), rot, atol=1e-6)


def test_left_handed_direction_uses_qfac(tmp_path):
    d = np.diag([1.0, 1.0, -1.0])
    path = str(tmp_path / 'flip.nii.gz')
    nifti.write_nifti(path, np.zeros((2, 3, 4), dtype=np.uint8), direction=tuple(d.reshape(-1)))
    raw = gzip.open(path, 'rb').read()
    assert struct.unpack_from('<f', raw, 76)[0] == -1.0
    _, geo = nifti.read_nifti(path)
    assert np.allclose(np.reshape(geo['direction'], (3, 3)), d)


def test_async_write_hook_numpy_and_tensor_subjects(tmp_path):
    rng = np.random.default_rng(1)
    prob = rng.random((5, 8, 6, 2)).astype(np.float32)
    prob /= prob.sum(-1, keepdims=True)
    props = types.SimpleNamespace(spacing=(1.0, 1.0, 2.5), origin=(0.0, 5.0, -3.0), direction=(1, 0, 0, 0, 1, 0, 0, 0, 1))
    hook = nifti.AsyncNiftiWriteHook()
    context = types.SimpleNamespace(test_dir=str(tmp_path))
    for subject, data in (('Brats18_A', {'probabilities': prob, 'properties': props, 'subject': 'Brats18_A'}),
                          ('Brats18_B', {'probabilities': torch.from_numpy(prob), 'properties': props, 'subject': 'Brats18_B'})):
        hook.on_test_subject_end(types.SimpleNamespace(subject_index=0, subject_data=data), None, context)
    hook.on_test_end(None, context)
    assert len(hook.written) == 4
    for subject in ('Brats18_A', 'Brats18_B'):
        p, geo = nifti.read_nifti(str(tmp_path / (subject + '_probabilities.nii.gz')))
        d, _ = nifti.read_nifti(str(tmp_path / (subject + '_prediction.nii.gz')))
        assert p.dtype == np.float32 and np.array_equal(p, prob[..., 1])               # foreground class, like WriteHook
        assert d.dtype == np.uint8 and np.array_equal(d, np.argmax(prob, -1).astype(np.uint8))
        assert np.allclose(geo['spacing'], props.spacing) and np.allclose(geo['origin'], props.origin)
