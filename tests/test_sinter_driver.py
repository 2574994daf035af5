"""The sinter file-protocol batch driver (ldpc_b200/sinter_decoder.py): b8 I/O and DEM parsing on the CPU, the
end-to-end decode_via_files against the oracle's per-shot BP + OSD-0 on the GPU."""
import numpy as np
import pytest
import scipy.sparse as sp

from ldpc_b200.sinter_decoder import DemMatrices, dem_text_to_matrices, read_b8, write_b8

# A hand-written detector error model in the style of a distance-5 repetition-code memory experiment (2 rounds),
# with a repeat block, shift_detectors, a duplicated mechanism and a hyperedge with a decomposition separator.
DEM = """
error(0.04) D0 L0
error(0.03) D0 D1
error(0.03) D1 D2
error(0.03) D2 D3
error(0.04) D3
error(0.02) D0 D4
error(0.02) D0 D4
error(0.05) D3 D2 L0
repeat 2 {
    error(0.01) D1 D5
    error(0.015) D2 D6 ^ D3 D7
    shift_detectors 4
}
detector(1, 0) D0
logical_observable L0
error(0.007) D1 D0
"""


def test_b8_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    for bits in (1, 7, 8, 9, 31, 64, 77):
        a = rng.integers(0, 2, size=(53, bits)).astype(np.uint8)
        path = tmp_path / f"x{bits}.b8"
        write_b8(path, a)
        assert path.stat().st_size == 53 * ((bits + 7) // 8)
        assert np.array_equal(read_b8(path, bits), a)
    # bit k of a shot is bit k % 8 of byte k // 8
    one = np.zeros((1, 12), np.uint8)
    one[0, 9] = 1
    write_b8(tmp_path / "one.b8", one)
    assert (tmp_path / "one.b8").read_bytes() == bytes([0, 2])


def test_dem_text_matches_reference_construction():
    mats = dem_text_to_matrices(DEM)
    H = mats.check_matrix.toarray()
    # distinct detector sets, in order of first appearance
    cols = [frozenset(np.nonzero(H[:, j])[0]) for j in range(H.shape[1])]
    want = [frozenset(s) for s in ([0], [0, 1], [1, 2], [2, 3], [3], [0, 4], [1, 5], [2, 3, 6, 7], [5, 9],
                                   [6, 7, 10, 11], [8, 9])]
    assert cols == want
    assert H.shape[0] == 12
    p = mats.priors
    assert p[0] == pytest.approx(0.04) and p[5] == pytest.approx(0.02 * 0.98 * 2)
    assert p[3] == pytest.approx(0.03 * 0.95 + 0.05 * 0.97)  # {D2, D3} seen twice, the second time as "D3 D2 L0"
    L = mats.observables_matrix.toarray()
    assert L.shape == (1, 11) and L[0, 0] == 1 and L[0, 3] == 1 and L[0].sum() == 2  # last observable set wins


@pytest.mark.gpu
def test_decode_via_files_matches_per_shot_oracle(tmp_path, port_oracle):
    from ldpc_b200.sinter_decoder import SinterBpOsdDecoder
    mats = dem_text_to_matrices(DEM)
    H = sp.csr_matrix(mats.check_matrix)
    rng = np.random.default_rng(3)
    shots = 6000
    err = (rng.random((shots, H.shape[1])) < mats.priors * 3).astype(np.uint8)
    dets = np.asarray((H.astype(np.int32) @ err.T.astype(np.int32)).T % 2).astype(np.uint8)
    dets[:50] = 0  # all-zero shots take the shortcut
    dem_path, in_path, out_path = tmp_path / "m.dem", tmp_path / "dets.b8", tmp_path / "obs.b8"
    dem_path.write_text(DEM)
    write_b8(in_path, dets)
    dec = SinterBpOsdDecoder(max_iter=10, bp_method="ms", ms_scaling_factor=0.625, osd_method="osd0")
    dec.decode_via_files(num_shots=shots, num_dets=H.shape[0], num_obs=1, dem_path=dem_path, dets_b8_in_path=in_path,
                         obs_predictions_b8_out_path=out_path, tmp_dir=tmp_path)
    got = read_b8(out_path, 1)
    # what the reference's loop computes per shot: BP, OSD-0 when BP fails, observables_matrix @ correction
    kw = dict(max_iter=10, bp_method="ms", ms_scaling_factor=0.625)
    r = port_oracle.decode_batch(H, dets, mats.priors, **kw)
    corr = r[0].copy()
    bad = ~r[1]
    if bad.any():
        corr[bad] = port_oracle.osd0_batch(H, dets[bad], r[3][bad])
    corr[~dets.any(axis=1)] = 0
    want = np.asarray((sp.csr_matrix(mats.observables_matrix, dtype=np.int32) @ corr.T.astype(np.int32)).T % 2)
    assert np.array_equal(got, want.astype(np.uint8))
    assert bad.any() and got.any()
    one = dec.decode(dets[100])
    assert np.array_equal(np.asarray(one).ravel() % 2, want[100])
