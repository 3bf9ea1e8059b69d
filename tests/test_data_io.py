"""Driver input formats: the reference's dense .npy layout, the sparse .npz form, header-less CSV, and the
block-diagonal replication fixture -- host logic, no GPU."""
import numpy as np
import pytest
import scipy.sparse as sp

from morfem_b200 import data_io, synthetic


def _model():
    ct, tt = synthetic.waveguide_operators(3, 2, 9)
    wp = synthetic.port_matrix(ct.shape[0], 2, 5)
    return sp.csc_array(ct), sp.csc_array(tt), sp.csc_array(wp)


def _same(a, b):
    return a.shape == b.shape and abs(sp.csc_array(a) - sp.csc_array(b)).max() == 0.0


@pytest.mark.parametrize("dense", [False, True])
def test_round_trip_sparse_npz_and_reference_npy(tmp_path, dense):
    ct, tt, wp = _model()
    data_io.save_operators(str(tmp_path), ct, tt, wp, dense=dense)
    got = data_io.load_operators(str(tmp_path))
    assert all(isinstance(g, sp.csc_array) for g in got)
    assert _same(got[0], ct) and _same(got[1], tt) and _same(got[2], wp)
    if dense:                                                    # exactly what main.py:21-23 reads
        assert np.array_equal(np.load(tmp_path / "Ct.npy"), ct.toarray())


def test_csv_like_convert_csv_to_json(tmp_path):
    ct, tt, wp = _model()
    for name, a in zip(data_io.NAMES, (ct, tt, wp)):
        np.savetxt(tmp_path / f"{name}.csv", a.toarray(), delimiter=",")      # header-less, like data_csv/*.csv
    got = data_io.load_operators(str(tmp_path))
    assert np.allclose(got[0].toarray(), ct.toarray(), rtol=1e-15, atol=0)
    assert got[2].shape == wp.shape


def test_missing_and_inconsistent_inputs(tmp_path):
    with pytest.raises(FileNotFoundError):
        data_io.load_operators(str(tmp_path))
    ct, tt, wp = _model()
    data_io.save_operators(str(tmp_path), ct, tt, wp[:-1])
    with pytest.raises(ValueError):
        data_io.load_operators(str(tmp_path))


def test_block_diagonal_replication_matches_the_dense_construction():
    ct, tt, wp = _model()
    k, n = 3, ct.shape[0]
    big_ct, big_tt, big_wp = data_io.replicate_block_diagonal(ct, tt, wp, k)
    dense = np.zeros((n * k, n * k))                              # the reference's loop, fake_interpolate_bigger_sample.py:5-8
    for i in range(k):
        dense[n * i:n * (i + 1), n * i:n * (i + 1)] = ct.toarray()
    assert np.array_equal(big_ct.toarray(), dense)
    assert big_tt.nnz == k * tt.nnz and _same(big_tt[n:2 * n, n:2 * n], tt)
    assert big_wp.shape == (n * k, wp.shape[1]) and _same(big_wp[2 * n:], wp)
    # the copies are uncoupled: the reduced problem of the replicated model has the spectrum of the single one, k times
    ev = np.linalg.eigvalsh(ct.toarray())
    assert np.allclose(np.linalg.eigvalsh(big_ct.toarray()), np.sort(np.repeat(ev, k)), atol=1e-9 * abs(ev).max())
    with pytest.raises(ValueError):
        data_io.replicate_block_diagonal(ct, tt, wp, 0)
