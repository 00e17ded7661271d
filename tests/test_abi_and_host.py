"""CPU-side checks of the product: the C-ABI library loads and exports every symbol that
include/amclj_b200.h declares, refuses to run without a GPU (no CPU fallback), and its
host-only planner agrees with the oracle's independent restatement."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import oracle
from amclj_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module", autouse=True)
def built():
    _lib.build()


def test_header_symbols_exported():
    hdr = (ROOT / "include" / "amclj_b200.h").read_text()
    declared = set(re.findall(r"\b(amclj_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"amclj_ctx", "amclj_params", "amclj_stats"}
    assert declared == set(_lib.SYMBOLS)
    L = C.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().amclj_abi_version() == 1


def test_params_struct_layout_matches_oracle_mirror():
    assert C.sizeof(_lib.Params) == C.sizeof(oracle.Params) == 112
    for preset in ("REF", "NS"):
        a, b = _lib.params(preset), oracle.params(preset)
        assert bytes(a) == bytes(b), preset


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.AmcljError) as e:
        _lib.Context(0)
    assert e.value.code == _lib.ERR_NODEVICE


def test_product_never_imports_oracle():
    for f in (ROOT / "amclj_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".h", ".cuh"):
            txt = f.read_text()
            assert "import oracle" not in txt and "from oracle" not in txt and "amclj_oracle" not in txt, f


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_plan_systematic_matches_oracle(world):
    rng = np.random.Generator(np.random.PCG64(world))
    for trial in range(50):
        n_global = int(rng.integers(1, 5000)) if trial % 2 else int(rng.integers(10**6, 10**7))
        totals = (rng.random(world) ** 4 * 2**44).astype(np.uint64)
        if trial % 7 == 0:
            totals[rng.integers(0, world)] = 0
        if totals.sum() == 0:
            totals[0] = 5
        u0 = float(rng.random())
        r, off, m_lo, m_cnt, send = _lib.plan_systematic(totals, n_global, u0)
        rr, roff, rlo, rcnt = oracle.plan(totals, n_global, u0)
        assert r == rr and np.array_equal(off, roff) and np.array_equal(m_lo, rlo) and np.array_equal(m_cnt, rcnt)
        assert m_cnt.sum() == n_global
        assert (send.sum(axis=1) == m_cnt).all()
        q, rem = divmod(n_global, world)
        assert (send.sum(axis=0) == [q + (1 if d < rem else 0) for d in range(world)]).all()


def test_plan_slots_agree_with_sequential_resampler():
    """The closed-form slot ranges reproduce the single-list systematic resampler."""
    rng = np.random.Generator(np.random.PCG64(3))
    n, world = 4001, 4
    w = (rng.random(n) ** 6).astype(np.float32)
    q, total = oracle.fixed_weights(w, oracle.wmax(w))
    u0 = 0.6180339
    ref = oracle.systematic(q, total, 0, oracle.systematic_offset(u0, total), n, 0, n)
    bounds = [0, 900, 2100, 2101, n]
    totals = np.array([q[bounds[g]:bounds[g + 1]].sum(dtype=np.uint64) for g in range(world)], np.uint64)
    r, off, m_lo, m_cnt, _ = _lib.plan_systematic(totals, n, u0)
    got = np.full(n, -1, np.int64)
    for g in range(world):
        loc = oracle.systematic(q[bounds[g]:bounds[g + 1]], total, int(off[g]), r, n, int(m_lo[g]), int(m_cnt[g]))
        assert (loc >= 0).all()
        got[m_lo[g]:m_lo[g] + m_cnt[g]] = loc + bounds[g]
    assert np.array_equal(got, ref)


def test_plan_rejects_bad_arguments():
    with pytest.raises(_lib.AmcljError):
        _lib.plan_systematic(np.array([1, 2], np.uint64), 10, 1.0)
