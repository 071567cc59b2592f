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


def _spec_header(nx, ny, nz, datatype, bitpix, qfac, pixdim, quatern, qoffset, srows):
    """A NIfTI-1 header built field by field from the offsets of nifti1.h (independent of nifti.header_bytes): every byte
    the standard does not require for a plain scanner-space volume stays zero."""
    h = bytearray(348)
    struct.pack_into('<i', h, 0, 348)                      # sizeof_hdr            @0
    struct.pack_into('<h', h, 40, 3)                       # dim[0]                @40
    struct.pack_into('<7h', h, 42, nx, ny, nz, 1, 1, 1, 1)  # dim[1..7]             @42
    struct.pack_into('<h', h, 70, datatype)                # datatype              @70
    struct.pack_into('<h', h, 72, bitpix)                  # bitpix                @72
    struct.pack_into('<f', h, 76, qfac)                    # pixdim[0] = qfac      @76
    struct.pack_into('<3f', h, 80, *pixdim)                # pixdim[1..3]          @80
    struct.pack_into('<f', h, 108, 352.0)                  # vox_offset            @108
    struct.pack_into('<f', h, 112, 1.0)                    # scl_slope             @112
    h[123] = 2                                             # xyzt_units = NIFTI_UNITS_MM  @123
    struct.pack_into('<h', h, 252, 1)                      # qform_code = NIFTI_XFORM_SCANNER_ANAT @252
    struct.pack_into('<h', h, 254, 1)                      # sform_code            @254
    struct.pack_into('<3f', h, 256, *quatern)              # quatern_b, c, d       @256
    struct.pack_into('<3f', h, 268, *qoffset)              # qoffset_x, y, z       @268
    for i, row in enumerate(srows):
        struct.pack_into('<4f', h, 280 + 16 * i, *row)     # srow_x, srow_y, srow_z @280, 296, 312
    h[344:348] = b'n+1\0'                                  # magic                 @344
    return bytes(h)


def test_brats_volume_header_is_byte_identical_to_the_hand_built_spec_header():
    """The geometry every BraTS-2018 volume carries in ITK terms (240 x 240 x 155 voxels of 1 mm, identity direction, origin
    (0, -239, 0) in LPS): NIfTI stores RAS, so x and y of origin and direction flip — a rotation of 180 degrees about z,
    quaternion (a, b, c, d) = (0, 0, 0, 1).  All 348 bytes are compared, then the 4-byte extension flag."""
    got = nifti.header_bytes((155, 240, 240), np.float32, spacing=(1.0, 1.0, 1.0), origin=(0.0, -239.0, 0.0))
    exp = _spec_header(240, 240, 155, 16, 32, 1.0, (1.0, 1.0, 1.0), (0.0, 0.0, 1.0), (-0.0, 239.0, 0.0),
                       ((-1.0, 0.0, 0.0, -0.0), (0.0, -1.0, 0.0, 239.0), (0.0, 0.0, 1.0, 0.0)))
    diff = [i for i in range(348) if got[i] != exp[i] and not (got[i] in (0x00, 0x80) and exp[i] in (0x00, 0x80) and i % 4 == 3)]   # -0.0 == 0.0
    assert not diff, ['@%d: %02x != %02x' % (i, got[i], exp[i]) for i in diff]
    assert got[348:352] == b'\0\0\0\0' and len(got) == 352
    # prediction volume: DT_UINT8
    got8 = nifti.header_bytes((155, 240, 240), np.uint8, origin=(0.0, -239.0, 0.0))
    assert struct.unpack_from('<hh', got8, 70) == (2, 8) and got8[:70] == got[:70] and got8[74:348] == got[74:348]


def test_independent_reader_recovers_the_itk_world_coordinates(tmp_path):
    """A reader written from the specification alone (method 3: sform; method 2: quaternion) maps voxel indices to RAS world
    coordinates; negating x and y must give ITK's physical points origin + direction @ (spacing * index) in LPS."""
    th = -0.4
    direction = np.array([[np.cos(th), 0.0, np.sin(th)], [0.0, 1.0, 0.0], [-np.sin(th), 0.0, np.cos(th)]])
    spacing, origin = np.array([0.8, 1.25, 2.0]), np.array([12.0, -30.5, 4.0])
    vol = np.arange(4 * 5 * 6, dtype=np.int16).reshape(4, 5, 6)
    path = str(tmp_path / 'geo.nii')
    nifti.write_nifti(path, vol, spacing=tuple(spacing), origin=tuple(origin), direction=tuple(direction.reshape(-1)))
    raw = open(path, 'rb').read()
    assert struct.unpack_from('<f', raw, 108)[0] == 352.0 and len(raw) == 352 + vol.nbytes
    data = np.frombuffer(raw, dtype='<i2', offset=352).reshape(4, 5, 6)                     # x fastest
    assert np.array_equal(data, vol)
    srow = np.array([struct.unpack_from('<4f', raw, 280 + 16 * i) for i in range(3)])
    qb, qc, qd = struct.unpack_from('<3f', raw, 256)
    qa = np.sqrt(max(0.0, 1.0 - qb * qb - qc * qc - qd * qd))
    rot = np.array([[qa * qa + qb * qb - qc * qc - qd * qd, 2 * qb * qc - 2 * qa * qd, 2 * qb * qd + 2 * qa * qc],
                    [2 * qb * qc + 2 * qa * qd, qa * qa + qc * qc - qb * qb - qd * qd, 2 * qc * qd - 2 * qa * qb],
                    [2 * qb * qd - 2 * qa * qc, 2 * qc * qd + 2 * qa * qb, qa * qa + qd * qd - qc * qc - qb * qb]])
    pixdim = np.array(struct.unpack_from('<4f', raw, 76))
    qoff = np.array(struct.unpack_from('<3f', raw, 268))
    for ijk in ((0, 0, 0), (5, 0, 0), (0, 4, 0), (0, 0, 3), (2, 3, 1)):
        idx = np.array(ijk, dtype=np.float64)
        itk_lps = origin + direction @ (spacing * idx)
        ras_s = srow[:, :3] @ idx + srow[:, 3]
        ras_q = rot @ (pixdim[1:4] * idx * np.array([1.0, 1.0, pixdim[0]])) + qoff
        for ras in (ras_s, ras_q):
            assert np.allclose(ras * np.array([-1.0, -1.0, 1.0]), itk_lps, atol=1e-4), (ijk, ras, itk_lps)
