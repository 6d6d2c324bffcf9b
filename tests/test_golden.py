"""Committed golden vectors (tests/golden/ref_vectors.npz, made by tests/golden/make_golden.py
from the REFERENCE's own C, src/chebpol.c + src/stalker.c) against

  * the C restatement in oracle/ (CPU, `-m "not gpu"`): bit-identical;
  * the CUDA kernels through the C ABI (`-m gpu`): STRICT kernels bit-identical (1e-14
    relative where libm exp/log/pow is involved), fast kernels within 1e-12 * max|f|.

Unlike tests/test_parity_gpu.py, nothing here needs a CPU oracle at run time: the expected
values are the stored outputs of the reference, so the GPU box checks the kernels against the
reference even though /root/reference does not exist there.
"""
import os

import numpy as np
import pytest

from chebpol_b200 import _lib

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz"))
TOL_REL = 1e-12


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same_bits(got, want):
    return np.array_equal(bits(got), bits(want))


def close(got, want, tol=TOL_REL, scale=None):
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got))
    s = np.abs(want[fin]).max() if scale is None else scale
    err = np.abs(got[fin] - want[fin]).max() if fin.any() else 0.0
    assert err <= tol * s, (err, s)


CHEB = ["cheb_a", "cheb_b", "cheb_c", "cheb_d"]
MGRID = [G[f"mlip_grid{d}"] for d in range(3)]
FGRID = [G[f"fh_grid{d}"] for d in range(3)]
FW = [G[f"fh_w{d}"] for d in range(3)]
SGRID = [G[f"stalk_grid{d}"] for d in range(2)]
POLYH_K = (3, 2, 1, -1.5)


def mlip_x(blend):
    return G["mlip_x"] if blend not in (1, 2) else G["mlip_x_inside"]


# ------------------------------------------------------------------ CPU: the restatement
def test_fixture_is_complete():
    assert len(G.files) == 88
    for k in G.files:
        assert G[k].dtype == np.float64


@pytest.mark.parametrize("name", CHEB)
def test_port_cheb(port, name):
    assert same_bits(port.evalcheb(G[name + "_coef"], G[name + "_x"]), G[name + "_y"])


def test_port_chebcoef(port):
    assert same_bits(port.chebcoef(G["coef_val"]), G["coef_out"])
    assert same_bits(port.chebcoef(G["coef_val"], dct=True), G["coef_dct"])


@pytest.mark.parametrize("blend", range(6))
def test_port_mlip(port, blend):
    assert same_bits(port.evalmlip(MGRID, G["mlip_values"], mlip_x(blend), blend=blend), G[f"mlip_y_blend{blend}"])


def test_port_fh(port):
    w = port.fhweights(FGRID, [3, 4, 4])
    for a, b in zip(w, FW):
        assert same_bits(a, b)
    assert same_bits(port.fh(G["fh_vals"], FGRID, FW, G["fh_x"]), G["fh_y"])


@pytest.mark.parametrize("k", POLYH_K)
def test_port_polyh(port, k):
    got = port.evalpolyh(G["polyh_knots"], G["polyh_w"], G["polyh_lw"], k, G["polyh_x"])
    assert same_bits(got, G[f"polyh_y_k{k}"])


def test_port_stalker_fit_and_eval(port):
    b, c, r = port.makestalk(G["stalk_val"], SGRID)
    assert same_bits(b, G["stalk_b"]) and same_bits(c, G["stalk_c"]) and same_bits(r, G["stalk_r"])
    a, hb, hc, hd = port.makehyp(G["stalk_val"], SGRID)
    assert same_bits(a, G["hyp_a"]) and same_bits(hb, G["hyp_b"]) and same_bits(hc, G["hyp_c"]) and same_bits(hd, G["hyp_d"])
    for blend in (0, 3, 4, 5):
        got = port.evalstalk(G["stalk_val"], SGRID, b, c, r, G["stalk_x"], blend=blend)
        assert same_bits(got, G[f"stalk_y_blend{blend}"])
        got = port.evalhyp(G["stalk_val"], SGRID, a, hb, hc, hd, G["stalk_x"], blend=blend)
        assert same_bits(got, G[f"hyp_y_blend{blend}"])


def sl_tables():
    dtri = np.asfortranarray(G["sl_dtri"].astype(np.int32))
    return G["sl_knots"], dtri, (G["sl_lumats"], G["sl_pivots"].astype(np.int32), G["sl_bbox"])


def test_port_simplexlinear(port):
    kn, dtri, ad = sl_tables()
    got = port.analyzesimplex(dtri, kn)
    assert all(np.array_equal(a, b) for a, b in zip(got, ad))
    for epol in (0, 1):
        for blend in (0, 3, 4, 5):
            y = port.evalsl(G["sl_x"], kn, dtri, ad, G["sl_val"], epol, blend)
            assert same_bits(y, G[f"sl_y_epol{epol}_blend{blend}"])


# ------------------------------------------------------------------ GPU: the CUDA kernels
def run(plan, x, kernel, variant=0):
    plan.set(_lib.OPT_KERNEL, kernel)
    plan.set(_lib.OPT_VARIANT, variant)
    return plan.eval_host(np.ascontiguousarray(x)), plan.kernel_used


@pytest.mark.gpu
@pytest.mark.parametrize("name", CHEB)
def test_gpu_cheb(gpu_ok, name):
    cf, x, want = G[name + "_coef"], G[name + "_x"], G[name + "_y"]
    plan = _lib.Plan.cheb(cf)
    got, k = run(plan, x, _lib.KERNEL_STRICT)
    assert k == "cheb_strict" and same_bits(got, want)
    # conditioning scale of a random tensor evaluated up to |x| = 1.2: sum|c| * prod max|T_k|
    scale = np.abs(cf).sum() * np.prod([np.cosh((n - 1) * np.arccosh(1.2)) for n in cf.shape])
    for variant in (_lib.VARIANT_MMA, _lib.VARIANT_SLAB, 0):
        got, k = run(plan, x, _lib.KERNEL_AUTO, variant)
        close(got, want, scale=scale)
    plan.close()


@pytest.mark.gpu
def test_gpu_chebcoef(gpu_ok):
    close(_lib.chebcoef(G["coef_val"]), G["coef_out"], tol=1e-13)
    close(_lib.chebcoef(G["coef_val"], dct=True), G["coef_dct"], tol=1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("blend", range(6))
def test_gpu_mlip(gpu_ok, blend):
    want = G[f"mlip_y_blend{blend}"]
    plan = _lib.Plan.mlip(MGRID, G["mlip_values"], blend)
    got, k = run(plan, mlip_x(blend), _lib.KERNEL_STRICT)
    assert k == "mlip_strict"
    if blend in (1, 2):  # libm exp on the host, device exp here
        close(got, want, tol=1e-14)
    else:
        assert same_bits(got, want)
    for variant in (1, 2):  # pair-gather kernel, cell-major kernel
        got, k = run(plan, mlip_x(blend), _lib.KERNEL_FAST, variant)
        assert k in ("mlip_pair", "mlip_cell"), k
        close(got, want)
    plan.close()


@pytest.mark.gpu
def test_gpu_fh(gpu_ok):
    for a, b in zip(_lib.fhweights(FGRID, [3, 4, 4]), FW):
        assert same_bits(a, b)
    plan = _lib.Plan.fh(G["fh_vals"], FGRID, FW)
    got, k = run(plan, G["fh_x"], _lib.KERNEL_STRICT)
    assert k == "fh_strict" and same_bits(got, G["fh_y"])
    got, k = run(plan, G["fh_x"], _lib.KERNEL_FAST)
    assert k == "fh_mma"
    # random values on a cubed grid: the Lebesgue constant of that dimension enters
    close(got, G["fh_y"], tol=1e-11)
    plan.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", POLYH_K)
def test_gpu_polyh(gpu_ok, k):
    want = G[f"polyh_y_k{k}"]
    plan = _lib.Plan.polyh(G["polyh_knots"], G["polyh_w"], G["polyh_lw"], k)
    got, used = run(plan, G["polyh_x"], _lib.KERNEL_STRICT)
    assert used == "polyh_strict"
    scale = np.abs(G["polyh_w"]).sum() * 8.0  # sum |w_i| * max phi over the box
    if k in (3, 1):
        assert same_bits(got, want)
    else:
        close(got, want, tol=1e-14, scale=scale)
    got, used = run(plan, G["polyh_x"], _lib.KERNEL_FAST)
    assert used == "polyh_tile"
    close(got, want, scale=scale)
    plan.close()


@pytest.mark.gpu
@pytest.mark.parametrize("blend", (0, 3, 4, 5))
def test_gpu_stalkers(gpu_ok, blend):
    tabs = (G["stalk_b"], G["stalk_c"], G["stalk_r"])
    plan = _lib.Plan.stalk(0, G["stalk_val"], SGRID, tabs, blend)
    want = G[f"stalk_y_blend{blend}"]
    got, k = run(plan, G["stalk_x"], _lib.KERNEL_STRICT)
    assert k == "stalk_strict"
    close(got, want, tol=1e-14)  # pow()
    got, k = run(plan, G["stalk_x"], _lib.KERNEL_FAST)
    assert k == "stalk_cell"
    close(got, want)
    plan.close()
    tabs = (G["hyp_a"], G["hyp_b"], G["hyp_c"], G["hyp_d"])
    plan = _lib.Plan.stalk(1, G["stalk_val"], SGRID, tabs, blend)
    want = G[f"hyp_y_blend{blend}"]
    got, k = run(plan, G["stalk_x"], _lib.KERNEL_STRICT)
    assert k == "stalk_strict" and same_bits(got, want)
    got, k = run(plan, G["stalk_x"], _lib.KERNEL_FAST)
    assert k == "stalk_cell"
    close(got, want)
    plan.close()


@pytest.mark.gpu
def test_gpu_simplexlinear(gpu_ok):
    kn, dtri, ad = sl_tables()
    assert all(np.array_equal(a, b) for a, b in zip(_lib.analyzesimplex(dtri, kn), ad))
    plan = _lib.Plan.sl(kn, dtri, ad, G["sl_val"])
    for epol in (0, 1):
        plan.set(_lib.OPT_EPOL, epol)
        for blend in (0, 3, 4, 5):
            plan.set(_lib.OPT_BLEND, blend)
            want = G[f"sl_y_epol{epol}_blend{blend}"]
            got, k = run(plan, G["sl_x"], _lib.KERNEL_STRICT)
            assert k == "sl_strict" and same_bits(got, want)
            got, k = run(plan, G["sl_x"], _lib.KERNEL_FAST)
            assert k == "sl_cell" and same_bits(got, want)
    plan.close()


@pytest.mark.gpu
def test_gpu_oldstalker(gpu_ok):
    import chebpol_b200 as cb

    grid = [G["oldst_grid0"], G["oldst_grid1"]]
    for tag, r in (("r2", [2.0, 2.0]), ("r15", [1.5, 1.0]), ("r0", [0.0, 2.0])):
        ip = cb.ipol(G["oldst_val"], grid=grid, k=r, method="oldstalker")
        for blend, name in ((0, "linear"), (3, "cubic")):
            want = G[f"oldst_y_{tag}_blend{blend}"]
            got = ip(G["oldst_x"].T, blend=name)
            assert ip.plan().kernel_used == "oldstalk"
            close(got, want, tol=1e-12)
        ip.close()
