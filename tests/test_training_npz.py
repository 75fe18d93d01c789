"""kc_training_write_npz (csrc/npzwrite.cpp): the reference's training-data file, TrainingWriteBuffers::writeToZipFile
(cpp/dataio/trainingwrite.cpp:566-587) with NumpyBuffer's headers (cpp/dataio/numpywrite.cpp:110-222).  Host-only."""
import zipfile

import numpy as np
import pytest


def _rows(rng, W, H, n):
    HW = W * H
    return dict(binaryInputNCHWPacked=rng.integers(0, 256, (n, 15, (HW + 7) // 8), dtype=np.uint8), globalInputNC=rng.random((n, 1), dtype=np.float32),
                policyTargetsNCMove=rng.integers(-5, 800, (n, 2, 4 * HW)).astype(np.int16), globalTargetsNC=rng.random((n, 64), dtype=np.float32),
                valueTargetsNCHW=rng.integers(-1, 2, (n, 5, H, W)).astype(np.int8))


@pytest.mark.parametrize("W,H,n", [(5, 5, 37), (6, 6, 2500), (7, 4, 3), (5, 5, 0)])
def test_npz_members_headers_and_values(built_lib, tmp_path, W, H, n):
    from katacoffee_b200 import backend
    rows = _rows(np.random.default_rng(n), W, H, n)
    path = str(tmp_path / "rows.npz")
    backend.writeTrainingNpz(path, W, H, rows)
    z = zipfile.ZipFile(path)
    assert z.testzip() is None                                            # CRCs
    assert z.namelist() == list(rows)                                     # member names and order of writeToZipFile, no extension
    assert all(i.compress_type == zipfile.ZIP_DEFLATED for i in z.infolist())
    descr = {"binaryInputNCHWPacked": "|u1", "globalInputNC": "<f4", "policyTargetsNCMove": "<i2", "globalTargetsNC": "<f4", "valueTargetsNCHW": "|i1"}
    for k, v in rows.items():
        raw = z.read(k)
        shape = ",".join(str(d) for d in v.shape)
        head = b"\x93NUMPY\x01\x00\xf6\x00" + ("{'descr':'%s','fortran_order':False,'shape':(%s)}" % (descr[k], shape)).encode()
        assert raw[:len(head)] == head and raw[len(head):255] == b" " * (255 - len(head)) and raw[255:256] == b"\n"   # numpywrite.cpp:133-222
        assert raw[256:] == v.tobytes()
    d = np.load(path)                                                     # what python/shuffle.py does
    for k, v in rows.items():
        assert d[k].dtype == v.dtype and d[k].shape == v.shape and (d[k] == v).all(), k
    assert not (tmp_path / "rows.npz.tmp").exists()


def test_npz_errors(built_lib, tmp_path):
    from katacoffee_b200 import backend, capi
    rows = _rows(np.random.default_rng(1), 5, 5, 4)
    with pytest.raises(capi.KCError, match="cannot open"):
        backend.writeTrainingNpz(str(tmp_path / "no_such_dir" / "rows.npz"), 5, 5, rows)
    with pytest.raises(AssertionError):
        backend.writeTrainingNpz(str(tmp_path / "rows.npz"), 6, 6, rows)   # shapes must match the board
