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
    assert _lib.lib().amclj_abi_version() == 2


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


def _end_cell_offsets(rng, n, r_m, res, span_cells):
    """(x1 - x0, y1 - y0) exactly as the kernels compute it: fp32 end point (one rounding, as the fma),
    IEEE fp32 divide by the resolution, floor(q + 0.5) — sensor.clj:99-103, :134-135, map.clj:36-46."""
    res32 = np.float32(res)
    ox = (rng.uniform(0, span_cells, n) * res).astype(np.float32)
    oy = (rng.uniform(0, span_cells, n) * res).astype(np.float32)
    # a share of the starts sits on the rounding boundary of a cell, where both ends flip most easily
    k = n // 4
    ox[:k] = ((np.floor(rng.uniform(0, span_cells, k)) + 0.5) * res).astype(np.float32)
    oy[:k] = ((np.floor(rng.uniform(0, span_cells, k)) + 0.5) * res).astype(np.float32)
    a = rng.uniform(-np.pi, np.pi, n)
    a[: n // 8] = np.round(a[: n // 8] / (np.pi / 4)) * (np.pi / 4)  # axis-aligned and diagonal beams
    c, s = np.cos(a).astype(np.float32), np.sin(a).astype(np.float32)
    R = np.float32(r_m)
    ex_w = (R.astype(np.float64) * c.astype(np.float64) + ox.astype(np.float64)).astype(np.float32)
    ey_w = (R.astype(np.float64) * s.astype(np.float64) + oy.astype(np.float64)).astype(np.float32)

    def cell(v):
        return np.floor((v / res32).astype(np.float32) + np.float32(0.5)).astype(np.int64)

    return cell(ex_w) - cell(ox), cell(ey_w) - cell(oy)


@pytest.mark.parametrize("r_m,res,span", [(5.6, 0.05, 700), (-1.0, 0.05, 700), (30.0, 0.05, 8192), (0.02, 0.05, 50),
                                          (0.3, 0.25, 9), (11.0, 1.0, 64), (100.0, 0.005, 300), (2.5, 0.1, 4000)])
def test_ray_ring_covers_every_end_cell(r_m, res, span):
    """The ring the ray tables (and the walk kernel's set-up records) are laid out on contains every
    end-cell offset the rounding of sensor.clj:134-135 can produce (a miss would only cost speed —
    the kernels walk such a ray — but the claim is that it never happens)."""
    rng = np.random.Generator(np.random.PCG64(int(abs(r_m) * 1000) + span))
    r_cells = abs(float(np.float32(r_m))) / float(np.float32(res))
    E, lo, cnt, n_entries = _lib.ray_ring(r_cells)
    assert n_entries > 0 and len(lo) == 2 * E + 1
    ex, ey = _end_cell_offsets(rng, 400_000, r_m, res, span)
    row = ey + E
    assert row.min() >= 0 and row.max() <= 2 * E
    t = np.abs(ex) - lo[row]
    assert np.all((t >= 0) & (t < cnt[row])), f"{np.count_nonzero((t < 0) | (t >= cnt[row]))} end cells off the ring"
    # and it is thin: about 2.7 cells per unit of circumference (plus the unused / sentinel slots)
    assert n_entries <= 2 * np.pi * (r_cells + 2) * 3.2 + 40


def test_ray_ring_limits():
    assert _lib.ray_ring(40_000.0)[3] == 0          # offsets would not fit 16 bits
    E, lo, cnt, n = _lib.ray_ring(0.0)               # zero-length rays: end = start, +-1 cell of rounding
    assert E == 2 and n > 0 and cnt[E] >= 2


def test_struct_layouts_match_the_c_header(tmp_path):
    """The ctypes mirrors of amclj_params / amclj_stats against what a C compiler makes of
    include/amclj_b200.h (sizeof and every field offset): the header is what a JNA / cgo binding reads."""
    import subprocess

    def fields(struct):
        return [n for n, _ in struct._fields_]

    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "amclj_b200.h"', "int main(void) {",
             '  printf("params %zu\\n", sizeof(amclj_params));', '  printf("stats %zu\\n", sizeof(amclj_stats));']
    for cname, struct in (("amclj_params", _lib.Params), ("amclj_stats", _lib.Stats)):
        for f in fields(struct):
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    assert int(out["params"]) == C.sizeof(_lib.Params) and int(out["stats"]) == C.sizeof(_lib.Stats)
    for cname, struct in (("amclj_params", _lib.Params), ("amclj_stats", _lib.Stats)):
        for f in fields(struct):
            assert int(out[f"{cname}.{f}"]) == getattr(struct, f).offset, f"{cname}.{f}"


def test_clojure_shim_reads_and_names_only_exported_symbols():
    """shim/src/amclj/gpu.clj cannot be executed here (no JVM), but it must at least be well-formed Clojure
    (read with the evaluator's reader) and every entry point it names must be one the library exports."""
    import re
    import sys
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import clj_eval
    text = (ROOT / "shim" / "src" / "amclj" / "gpu.clj").read_text()
    forms = clj_eval.read_all(text)
    assert len(forms) > 20 and forms[0][0] == "ns"
    named = set(re.findall(r'"(amclj_[a-z_0-9]+)"', text)) - {"amclj_b200"}  # the library's own name
    assert len(named) >= 15
    L = _lib.lib()
    missing = [s for s in named if not hasattr(L, s)]
    assert not missing, missing
    patch = (ROOT / "shim" / "pf.clj.patch").read_text()
    assert "apply-motion-model" in patch and "roulette-selection" in patch
