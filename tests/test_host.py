"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header
declares; argument validation that precedes device selection mirrors the reference's
entry checks; the package fails loudly -- never falls back -- without a GPU; the
host mirror of the R interface coerces arguments as R/tools.R:9-20 does; and the
multi-rank point partition tiles the batch (world_size-2 gloo run)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import chebpol_b200 as cb
from chebpol_b200 import _lib, shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "chebpol_cuda.h")).read()
    return sorted(set(re.findall(r"CHEBPOL_CUDA_API[^;(]*?\b(chebpol_cuda_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert getattr(L, s) is not None, s
    assert sorted(_lib.EXPORTS) == syms
    # and nothing from the oracle is linked into the product
    out = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle_" not in out and "ref_evalcheb" not in out


def test_r_glue_compiles_against_shim():
    """The .Call wrappers (csrc/r_glue.c) type-check against the R API declarations of oracle/shim."""
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "chebpol_b200", "csrc"), "rglue-check"], capture_output=True, text=True)
    assert r.returncode == 0 and "r_glue.c: ok" in r.stdout, r.stdout + r.stderr


def test_strerror_and_version():
    L = _lib.lib()
    assert L.chebpol_cuda_version() >= 100
    assert L.chebpol_cuda_strerror(0) == b"success"
    for code in range(1, 9):
        assert len(L.chebpol_cuda_strerror(code)) > 0
    assert b"unknown" in L.chebpol_cuda_strerror(99)


def test_argument_validation_precedes_device():
    """Mirrors 'rank must be positive' (src/chebpol.c:352,232) and the size checks."""
    L = _lib.lib()
    h = C.c_void_p()
    cf = np.zeros(4)
    d = np.array([2, 2], dtype=np.int32)
    dp = d.ctypes.data_as(C.POINTER(C.c_int))
    assert L.chebpol_cuda_plan_cheb(cf.ctypes.data, dp, 0, None, None, None, 0, C.byref(h)) == 2
    assert b"rank must be positive" in L.chebpol_cuda_last_error()
    assert L.chebpol_cuda_plan_cheb(cf.ctypes.data, dp, 40, None, None, None, 0, C.byref(h)) == 2
    assert L.chebpol_cuda_plan_cheb(None, dp, 2, None, None, None, 0, C.byref(h)) == 1
    bad = np.array([2, 0], dtype=np.int32)
    assert L.chebpol_cuda_plan_cheb(cf.ctypes.data, bad.ctypes.data_as(C.POINTER(C.c_int)), 2, None, None, None, 0,
                                    C.byref(h)) == 3
    huge = np.array([65536, 65536], dtype=np.int32)
    assert L.chebpol_cuda_plan_cheb(cf.ctypes.data, huge.ctypes.data_as(C.POINTER(C.c_int)), 2, None, None, None, 0,
                                    C.byref(h)) == 3
    mid = np.zeros(2)
    assert L.chebpol_cuda_plan_cheb(cf.ctypes.data, dp, 2, mid.ctypes.data, None, None, 0, C.byref(h)) == 1
    g = [np.linspace(0, 1, 2), np.linspace(0, 1, 2)]
    gl, keep = _lib.ptr_list(g)
    assert L.chebpol_cuda_plan_mlip(gl, dp, 2, cf.ctypes.data, 9, 0, C.byref(h)) == 4  # EBLEND
    assert L.chebpol_cuda_plan_eval_host(None, None, 1, None) == 1


@pytest.mark.skipif(_lib.device_count() > 0, reason="a GPU is visible")
def test_argument_validation_of_the_widened_entry_points():
    """general / simplex-linear / oldstalker: argument checks run before any device is touched, with the
    reference's messages where it has one (src/stalker.c:871-874)."""
    from chebpol_b200 import interp

    # general: spline knots must be strictly increasing, one table per dimension
    g = np.linspace(0, 1, 5) ** 2
    tab = interp.hyman_spline(g, interp.chebknots([5])[0])
    bad = (tab[0][::-1].copy(),) + tuple(tab[1:])
    with pytest.raises(_lib.ChebpolCudaError) as e:
        _lib.Plan.chebg(np.zeros(5), [bad])
    assert e.value.code == 1 and "strictly increasing" in str(e.value)
    with pytest.raises(ValueError):
        _lib.Plan.chebg(np.zeros((5, 5)), [tab])
    # simplex-linear: knot indices of the triangulation are checked
    knots = np.array([[0.0, 1.0, 0.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    dtri = np.array([[1, 2], [2, 3], [3, 5]], dtype=np.int32)  # 5 is not a knot
    ad = interp.analyzesimplex(np.array([[1, 2], [2, 3], [3, 4]], dtype=np.int32), knots)
    with pytest.raises(_lib.ChebpolCudaError) as e:
        _lib.Plan.sl(knots, dtri, ad, np.zeros(4))
    assert e.value.code == 1 and "not a knot index" in str(e.value)
    with pytest.raises(ValueError):
        _lib.Plan.sl(knots, dtri[:2], ad, np.zeros(4))  # dtri must have dim+1 rows
    with pytest.raises(_lib.ChebpolCudaError) as e:
        _lib.evalsl(np.zeros((1, 2)), knots, np.array([[1, 2], [2, 3], [3, 4]], dtype=np.int32), ad, np.zeros(4), blend=9)
    assert e.value.code == 4  # EBLEND
    # oldstalker: NaN degree on a non-uniform grid is refused like the reference built without GSL
    g2 = [np.array([0.0, 0.1, 0.4, 1.0])]
    with pytest.raises(_lib.ChebpolCudaError) as e:
        _lib.Plan.oldstalk(np.zeros(4), g2, [np.nan], interp.stalkercontext(g2, [2.0]), [0])
    assert "not supported without GSL" in str(e.value)
    with pytest.raises(_lib.ChebpolCudaError) as e:
        _lib.Plan.oldstalk(np.zeros(1), [np.array([0.0])], [2.0], interp.stalkercontext([np.array([0.0, 1.0])], [2.0])[:0] or
                           ([np.zeros(1)], [np.zeros(1)], [np.zeros(1)]), [1])
    assert e.value.code == 3  # a stalker grid needs at least two knots


def test_no_cpu_fallback_without_gpu():
    cf = np.ones((3, 3))
    with pytest.raises(cb.ChebpolCudaError) as e:
        _lib.Plan.cheb(cf)
    assert e.value.code == 5  # ENODEV
    ip = cb.chebappx(np.ones((3, 3)))
    with pytest.raises(cb.ChebpolCudaError):
        ip(np.array([0.1, 0.2]))
    with pytest.raises(cb.ChebpolCudaError):
        _lib.evalcheb(cf, np.zeros((4, 2)))


def test_missing_library_is_loud(tmp_path, monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.lib()


def test_vectorfun_coercion_rules():
    """R/tools.R:9-20 via Interpolant._coerce (no device needed)."""
    ip = cb.Interpolant("chebyshev", 3, None, coef=np.zeros((2, 2, 2)))
    x, _ = ip._coerce(np.arange(6.0).reshape(3, 2))  # K x P matrix
    assert x.shape == (2, 3) and np.array_equal(x[1], [1, 3, 5])
    x, single = ip._coerce([1.0, 2.0, 3.0])
    assert x.shape == (1, 3) and single
    with pytest.warns(UserWarning):
        x, _ = ip._coerce(np.arange(6.0))
    assert x.shape == (2, 3) and np.array_equal(x[0], [0, 1, 2])  # dim(x) <- c(arity, numvec)
    with pytest.raises(ValueError):
        ip._coerce(np.arange(4.0))
    with pytest.raises(ValueError):
        ip._coerce(np.zeros((2, 5)))
    ip1 = cb.Interpolant("chebyshev", 1, None, coef=np.zeros(4))
    x, _ = ip1._coerce(np.arange(5.0))
    assert x.shape == (5, 1)


def test_ipol_front_end_checks():
    with pytest.raises(ValueError):
        cb.ipol(lambda x: 0.0, method="chebyshev")  # dims must be specified
    with pytest.raises(ValueError):
        cb.ipol(np.zeros(4), method="multilinear")  # grid must be specified
    with pytest.raises(ValueError):
        cb.ipol(np.zeros(3), grid=[np.array([0.0, 2.0, 1.0])], method="fh")  # unsorted
    with pytest.raises(NotImplementedError):
        cb.ipol(np.zeros(4), grid=[np.linspace(0, 1, 4)], method="crbf")  # disabled in the reference too
    with pytest.raises(ValueError):
        cb.ipol(np.zeros(4), method="simplexlinear")  # knots must be specified
    sl = cb.ipol(lambda x: float(np.sum(x)), knots=np.random.default_rng(0).uniform(0, 1, (2, 30)),
                 method="simplexlinear", fit="host")
    assert sl.kind == "simplexlinear" and sl.arity == 2 and sl.t["dtri"].shape[0] == 3 and len(sl.t["adata"]) == 3
    with pytest.raises(ValueError):
        cb.ipol(np.zeros(4), method="hstalker")  # grid must be specified
    hs = cb.ipol(lambda x: float(np.sum(x ** 2)), grid=[np.linspace(0, 1, 5)] * 2, method="hstalker")
    assert hs.kind == "hstalker" and hs.t["tables"][1].shape == (2, 25) and hs.t["tables"][0].shape == (25,)
    with pytest.raises(ValueError):
        cb.ipol(np.zeros(4), method="polyharmonic")  # Must specify knots
    ip = cb.ipol(np.zeros(9), grid=[np.linspace(0, 1, 3)] * 2, method="multi")  # partial match
    assert ip.kind == "multilinear" and ip.arity == 2
    ip = cb.ipol(np.arange(24.0), dims=[2, 3, 4], method="cheb")
    assert ip.t["coef"].shape == (2, 3, 4)
    with pytest.raises(ValueError):
        cb.ipol(np.zeros(4), method="general")  # grid must be specified
    g = cb.ipol(lambda x: float(np.sum(x)), grid=[np.linspace(0, 1, 5) ** 2, np.linspace(0, 2, 4)], method="general")
    assert g.kind == "general" and g.arity == 2 and g.t["coef"].shape == (5, 4) and len(g.t["splines"]) == 2
    assert all(len(t) == 5 for t in g.t["splines"])
    ip = cb.ipol(np.zeros((5, 5)), grid=[np.linspace(0, 1, 5)[::-1], np.linspace(0, 1, 5)], method="fh")
    assert [len(w) for w in ip.t["weights"]] == [5, 5]  # decreasing grids are legal (R/ipol.R:241-243)
    ip = cb.ipol(np.zeros(15), method="uniform")
    assert ip.kind == "uniform" and list(ip.t["ugm_n"]) == [15]


def test_polyh_fit_mirrors_the_reference_system(port):
    """polyh.real (R/polyharmonic.R:48-131): the fitted weights solve the saddle system, so the
    spline (evaluated by the CPU oracle here; no GPU in this test) reproduces the values on the
    knots, and the side conditions sum w = 0, knots @ w = 0 hold."""
    rng = np.random.default_rng(5)
    knots = rng.uniform(0, 1, (3, 60))
    f = lambda x: 10.0 / (1.0 + 25.0 * np.mean(x ** 2))
    for k in (1, 2, 3, -2.0):
        ip = cb.polyh(f, knots, k=k)
        t = ip.t
        assert ip.kind == "polyharmonic" and ip.arity == 3 and t["norm_a"] is None
        got = port.evalpolyh(t["knots"], t["weights"], t["lweights"], t["k"], knots.T)
        want = np.array([f(knots[:, i]) for i in range(60)])
        assert np.abs(got - want).max() < 1e-8
        assert abs(t["weights"].sum()) < 1e-8 and np.abs(t["knots"] @ t["weights"]).max() < 1e-8
    # knots outside [0,1]: normalised to the unit cube, the map is carried for the kernel
    ip = cb.polyh(f, 20 * knots - 3, k=3)
    assert np.allclose(ip.t["norm_b"], (20 * knots - 3).min(axis=1)) and ip.t["knots"].min() == 0.0
    assert ip.t["knots"].max() == 1.0


def test_stalker_fits_mirror_the_reference(ref):
    """makehyp / makestalk (src/stalker.c:446-535, 271-364, no GSL) in numpy against the reference C:
    hyperbolic tables identical, stalker tables to an ulp of pow; NaN (parabola) pattern equal."""
    import warnings

    rng = np.random.default_rng(0)
    for dims, uniform in [((9,), True), ((8, 7), False), ((6, 5, 7), True), ((5, 6, 4, 5), False), ((2, 3), True)]:
        grid = [np.linspace(-1, 1, n) if uniform else np.sort(rng.uniform(-1, 1, n)) for n in dims]
        val = rng.standard_normal(dims)
        if len(dims) == 1:
            val[3:6] = 1.0       # constant triple
            val[6:9] = [0, 1, 2]  # linear triple
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ts = cb.interp.makestalk(val, grid)
        th = cb.interp.makehyp(val, grid)
        for mine, theirs in zip(th, ref.makehyp(val, grid)):
            assert np.array_equal(np.isnan(mine), np.isnan(theirs))
            assert np.array_equal(np.nan_to_num(mine), np.nan_to_num(theirs))
        for mine, theirs in zip(ts, ref.makestalk(val, grid)):
            assert np.allclose(mine, theirs, rtol=1e-14, atol=1e-15)
    # the parabola marker: symmetric values around an interior extremum on a non-uniform grid
    g = [np.array([0.0, 1.0, 3.0])]
    th = cb.interp.makehyp(np.array([1.0, 0.0, 4.0]), g)
    assert np.isnan(th[2][0, 1]) and np.isnan(ref.makehyp(np.array([1.0, 0.0, 4.0]), g)[2][0, 1])


def test_shard_range_tiles_the_batch():
    for npts in [0, 1, 7, 1000, 10**9 + 7]:
        for world in [1, 2, 3, 4, 8]:
            prev = 0
            for r in range(world):
                lo, hi = shard.shard_range(npts, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == npts
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from chebpol_b200 import shard
rank, local, world = shard.init_process_group("gloo")
assert world == 2
npts = 1001
lo, hi = shard.shard_range(npts, rank, world)
# every rank owns a slice; the slices' sizes sum to P; MAX reduction picks the slower rank
tot = shard.sum_over_ranks(hi - lo)
assert tot == npts, tot
shard.barrier()
mx = shard.max_over_ranks(10.0 + rank)
assert mx == 11.0, mx
# a replicated "tensor" evaluated on each rank's slice, gathered: equals the whole
x = np.linspace(-1, 1, npts)
mine = torch.zeros(npts, dtype=torch.float64)
mine[lo:hi] = torch.from_numpy(np.polynomial.chebyshev.chebval(x[lo:hi], [1.0, 0.5, 0.25]))
dist.all_reduce(mine)
assert np.allclose(mine.numpy(), np.polynomial.chebyshev.chebval(x, [1.0, 0.5, 0.25]))
dist.destroy_process_group()
open(os.path.join({out!r}, "ok_%d" % rank), "w").write("ok")
"""


def test_two_rank_gloo_partition(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER.format(root=ROOT, out=str(tmp_path)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "ok_0").exists() and (tmp_path / "ok_1").exists(), r.stdout + r.stderr


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints ONE JSON line with
    the contract's keys, needs no GPU, and is silent on non-zero ranks."""
    import json
    import sys

    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1",
           "--warmup", "1", "--ref-seconds", "0.3", "--workloads", "4", "--sub-cpu-seconds", "0.2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]
    # the other BASELINE configs ride in `workloads`, timed by the same procedure
    assert [w["cfg"] for w in d["workloads"]] == [4] and d["workloads"][0]["value"] > 0
    assert d["workloads"][0]["kind"] == d["cpu_baseline"]["kind"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_interpolants_pickle_like_saved_r_closures():
    """R interpolants survive save()/load() because they are plain closures over their tensors
    (R/chebpol-package.R:18-22, SURVEY 5 "checkpoint / resume"): the mirror pickles the host tensors and
    drops device plans, which are rebuilt on first use."""
    import pickle

    ip = cb.ipol(lambda x: float(np.sum(x)), grid=[np.linspace(0, 1, 5) ** 2, np.linspace(0, 2, 4)], method="general")
    ip._plans[0] = object()  # stands for a live device plan (a ctypes handle cannot be pickled)
    ip2 = pickle.loads(pickle.dumps(ip))
    assert ip2.kind == "general" and ip2._plans == {} and ip2.arity == 2
    assert np.array_equal(ip2.t["coef"], ip.t["coef"]) and len(ip2.t["splines"]) == 2
    ip._plans.clear()
    ml = pickle.loads(pickle.dumps(cb.ipol(np.arange(9.0), grid=[np.linspace(0, 1, 3)] * 2, method="multilinear")))
    assert ml.kind == "multilinear" and np.array_equal(ml.t["values"], np.arange(9.0))


def test_synthetic_points_are_partition_independent():
    """coordinate d of global point i depends on (seed, i, d) only, so any slice of a batch --
    another GPU's shard, the CPU baseline's prefix, a strided parity sample -- regenerates identical values."""
    from chebpol_b200 import synth

    whole = synth.points(0, 5000, 6)
    assert whole.shape == (5000, 6) and whole.flags.c_contiguous
    assert np.array_equal(synth.points(1234, 777, 6), whole[1234:2011])
    idx = np.arange(3, 5000, 97)
    assert np.array_equal(synth.points_at(idx, 6), whole[idx])
    assert not np.array_equal(synth.points(0, 100, 6, seed=43), whole[:100])
    assert -1.0 <= whole.min() and whole.max() < 1.0 and abs(whole.mean()) < 0.02
    big = synth.points(10**9 - 3, 3, 6)  # 64-bit element counters: cfg5's last points
    assert np.array_equal(big, synth.points_at([10**9 - 3, 10**9 - 2, 10**9 - 1], 6))
    lo_hi = synth.points(0, 1000, 2, lo=-1.2, hi=1.2)
    assert lo_hi.min() >= -1.2 and lo_hi.max() < 1.2


def test_fingerprint_is_thread_independent_and_sensitive():
    """The plan cache's 128-bit content fingerprint (csrc/hash128.h): the same value whatever the
    number of host threads, different for any single changed word, length and order sensitive."""
    rng = np.random.default_rng(5)
    a = rng.random(3 * (1 << 20) + 12345)  # > 8 MB: the threaded path; ragged last segment
    h1 = _lib.fingerprint(a, 1)
    assert _lib.fingerprint(a, 4) == h1 and _lib.fingerprint(a, 7) == h1
    for i in (0, 1, 131071, 131072, a.size // 2, a.size - 1):
        b = a.copy()
        b[i] = np.nextafter(b[i], 2.0)
        assert _lib.fingerprint(b, 4) != h1, i
    assert _lib.fingerprint(a[:-1], 1) != h1
    assert _lib.fingerprint(np.zeros(0), 1) != _lib.fingerprint(np.zeros(1), 1)
    small = rng.integers(0, 255, 1001).astype(np.uint8)  # unaligned tail, 4-byte aligned int arrays are legal input
    assert _lib.fingerprint(small, 1) != _lib.fingerprint(small[::-1].copy(), 1)
    assert _lib.fingerprint(small[1:], 1) == _lib.fingerprint(small[1:].copy(), 1)  # odd address, same bytes


def test_plan_cache_never_aliases_two_interpolants():
    """with every fingerprint forced equal, two different tensors of the same shape
    must still be told apart -- the cache compares the bytes of the retained host copy."""
    rng = np.random.default_rng(6)
    a = rng.random(4096)
    b = a.copy()
    b[2048] += 1e-9
    assert _lib.cache_would_hit(a, a.copy()) and not _lib.cache_would_hit(a, b)
    assert not _lib.cache_would_hit(a, a[:-1])
    _lib.cache_weak_hash(True)
    try:
        assert _lib.cache_would_hit(a, a.copy(), retain=True)
        assert not _lib.cache_would_hit(a, b, retain=True)       # bytes decide
        assert not _lib.cache_would_hit(a, a[:-1], retain=True)  # byte count is part of the key
        assert _lib.cache_would_hit(a, b, retain=False)          # what the hook removes: only the fingerprint separates entries above 32 MB
    finally:
        _lib.cache_weak_hash(False)
    assert not _lib.cache_would_hit(a, b, retain=False)


def test_sl_rejects_malformed_pivots(port):
    """pivots outside dgetrf's range would index out of bounds on the device; the check runs
    before any device is touched."""
    import cases

    kn, dtri, kv = cases.sl_case(2, 50)
    lu, piv, bb = port.analyzesimplex(dtri, kn)
    bad = np.array(piv, dtype=np.int32, copy=True)
    bad.flat[1] = 7
    with pytest.raises(_lib.ChebpolCudaError) as ei:
        _lib.Plan.sl(kn, dtri, (lu, bad, bb), kv)
    assert ei.value.code == 1 and "pivot" in str(ei.value)


def test_mma_engine_geometry_invariants():
    """Host-only layout logic of the DMMA contraction engine (csrc/contract_mma.cu), over a few hundred shapes:
    whatever it chooses must cover every tensor row exactly once, fit the shared memory of a CTA, keep the B-fragment
    loads conflict-free (row stride rule) and give every k of a row its own position."""
    import ctypes as C

    L = _lib.lib()
    L.chebpol_cuda_debug_mma_geometry.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int64)]
    L.chebpol_cuda_debug_mma_row_pos.argtypes = [C.c_int, C.c_int]
    rng = np.random.default_rng(7)
    shapes = [(10,) * 5, (12,) * 6, (40, 40, 40), (20, 20), (100, 100), (150, 3), (192,), (193,), (8,) * 6, (6,) * 7,
              (16,) * 4, (24,) * 3, (3, 3, 3, 3, 3), (1, 4), (4, 1, 3), (2,) * 7, (33, 5), (12, 12, 12, 3), (77, 2), (80, 2)]
    shapes += [tuple(int(v) for v in rng.integers(1, 41, int(rng.integers(1, 7)))) for _ in range(300)]
    taken = 0
    for kind in (0, 1):
        for dims in shapes:
            if np.prod(dims, dtype=np.float64) > 5e6:
                continue
            for npt in (32, 64, 96, 128, 256):
                out = (C.c_int64 * 16)()
                d = (C.c_int * len(dims))(*dims)
                assert L.chebpol_cuda_debug_mma_geometry(kind, d, len(dims), npt, out) == 0
                ok, nfold, K, kc, rs, rows, rps, nslabs, ns, stage, resident, pairs, smem, nA, nAp, nrp = (int(v) for v in out)
                if not ok:
                    continue
                taken += 1
                assert 1 <= nfold <= 3 and K == int(np.prod(dims[:nfold])) and K <= 192 and 4 * kc >= K
                assert len(dims) - nfold <= 4
                assert nA == (dims[nfold] if nfold < len(dims) else 1) and nAp % 2 == 0 and nA <= nAp <= nA + 1
                assert nrp == int(np.prod(dims[nfold + 1:])) and rows == -(-nrp // 8) * nAp * 8
                assert rs >= 4 * kc and rs % 16 == (8 if pairs else rs % 16) and (pairs or rs % 16 in (4, 12))
                assert pairs == (kc >= 20)
                assert rps % 16 == 0 or rps == rows
                assert (nslabs - 1) * rps < rows <= nslabs * rps            # the slabs cover every row once
                assert stage == rps * rs * 8 and ns >= (nslabs if resident else 2) and ns <= 8
                assert smem <= 220 * 1024                                   # ring + basis table inside the budget
                # every coefficient of a row has its own position inside the row's stride
                pos = [L.chebpol_cuda_debug_mma_row_pos(kc, k) for k in range(4 * kc)]
                assert sorted(pos) == list(range(4 * kc)) and max(pos) < rs
                if pairs:   # lane t of a quad finds chunks 2q and 2q+1 side by side
                    for q in range(kc // 2):
                        for t in range(4):
                            assert pos[4 * (2 * q) + t] == 8 * q + 2 * t and pos[4 * (2 * q + 1) + t] == 8 * q + 2 * t + 1
    assert taken > 500
