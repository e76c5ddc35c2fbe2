"""Post-fit reductions on the device and device-to-device data handles, through the C ABI."""
import numpy as np
import pytest

from resnmtf_b200 import _lib as L
from resnmtf_b200 import bicluster as B
from resnmtf_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from resnmtf_b200.device import Context

    c = Context()
    yield c
    c.close()


def factor_like_columns(n, m, rng):
    """Columns shaped like normalised F factors: most entries near zero, a planted block well above 1/n; one column
    of noise only, one constant column, one with an exact-zero tail."""
    cols = np.abs(rng.standard_normal((n, m))) * 0.05 / n
    for c in range(m - 3):
        rows = rng.random(n) < 0.15
        cols[rows, c] += (3.0 + rng.random(int(rows.sum()))) / n
    cols[:, m - 2] = 1.0 / n
    cols[n // 2:, m - 1] = 0.0
    return np.asfortranarray(cols / cols.sum(axis=0)[None, :])


@pytest.mark.parametrize("n", [700, 1024, 20000])
def test_jsd_pairs_match_the_host_restatement(ctx, n):
    """resnmtf_jsd_pairs against jsd_calc (R/utils.r:95-106 restated in bicluster.py: binning + FFT convolution +
    interpolation + JSD) on every ordered pair of 8 columns: 1e-9 relative (the kernel evaluates the convolution as
    direct sums instead of an FFT), bit-identical between two calls."""
    from resnmtf_b200.device import jsd_pairs

    rng = np.random.default_rng(100 + n)
    cols = factor_like_columns(n, 8, rng)
    bw = np.array([B.bw_nrd0(cols[:, c]) for c in range(8)])
    vmax = cols.max(axis=0)
    pa, pb = np.meshgrid(np.arange(8), np.arange(8), indexing="ij")
    pa, pb = pa.ravel(), pb.ravel()
    got = jsd_pairs(ctx, cols, bw, vmax, pa, pb)
    again = jsd_pairs(ctx, cols, bw, vmax, pa, pb)
    assert np.array_equal(got, again)
    want = np.array([B.jsd_calc(cols[:, a], cols[:, b]) for a, b in zip(pa, pb)])
    assert np.all(np.isfinite(want))
    assert np.allclose(got, want, rtol=1e-9, atol=1e-13), float(np.max(np.abs(got - want)))
    assert np.all(np.abs(got[pa == pb]) <= 1e-13)  # a column against itself


def test_jsd_pairs_argument_checks(ctx):
    from resnmtf_b200.device import jsd_pairs

    cols = factor_like_columns(64, 4, np.random.default_rng(0))
    bw = np.ones(4)
    with pytest.raises(L.ResnmtfError):
        jsd_pairs(ctx, cols, bw, cols.max(axis=0), [0, 4], [1, 2])  # column index out of range
    assert jsd_pairs(ctx, cols, bw, cols.max(axis=0), [], []).size == 0


def test_spurious_scores_on_the_device_equal_the_host_loops(ctx):
    """_jsd_scores_device (one launch per view) returns the threshold scores of calculate_f_shuffle_jsd in the
    reference's order and the per-cluster means of check_biclusters (R/obtain_bicl.r:55-68, 113-133)."""
    rng = np.random.default_rng(5)
    n, k, reps = 900, 3, 3
    f_main = factor_like_columns(n, k + 3, rng)[:, :k]
    f_mess = [[factor_like_columns(n, k + 3, rng)[:, :k]] for _ in range(reps)]
    thr, per_cluster = B._jsd_scores_device(f_mess, f_main, 0, reps, k, ctx)
    want_thr = []
    for j in range(reps - 1):
        want_thr += B.calculate_f_shuffle_jsd(f_mess, 0, j, reps, k)
    noise = np.concatenate([f_mess[r][0] for r in range(reps)], axis=1)
    want_pc = [np.mean([B.jsd_calc(f_main[:, c], noise[:, j]) for j in range(noise.shape[1])]) for c in range(k)]
    assert len(thr) == len(want_thr)
    assert np.allclose(thr, want_thr, rtol=1e-9, atol=1e-13)
    assert np.allclose(per_cluster, want_pc, rtol=1e-9, atol=1e-13)


def test_data_handle_from_a_device_pointer_equals_the_host_upload(ctx):
    """resnmtf_data_create_device (a view gathered on the GPU) gives the same fit as resnmtf_data_create."""
    import torch

    from resnmtf_b200.device import DeviceData, DeviceFit

    rng = np.random.default_rng(8)
    n, p, k = 1000, 333, 4
    x = synth.prep(synth.planted_view(n, p, 3, rng)[0])
    f0, s0, g0 = synth.random_factors(n, p, k, rng)
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).to(torch.device("cuda", ctx.device))
    torch.cuda.synchronize()
    outs = []
    for handle in (DeviceData(ctx, x), DeviceData.from_device(ctx, xt.data_ptr(), n, p)):
        with DeviceFit(ctx, [n], [p], [k]) as fit:
            fit.attach_data(0, handle)
            fit.set_factors(0, f0, s0, g0)
            fit.run(12)
            outs.append((fit.get_factors(0), fit.errors()))
        handle.close()
    for a, b in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(outs[0][1], outs[1][1])
