"""CPU-side checks of the drop-in boundary: the shared library builds for sm_100a, loads, exports every
symbol include/apertis_b200.h declares, and the ctypes table mirrors the header's arities.  No compute calls."""
import os
import re

from apertis_llm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "apertis_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    fns = {}
    for m in re.finditer(r"\b(int|size_t|int64_t)\s+(ab_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        fns[m.group(2)] = n
    return fns


def test_library_builds_and_exports_every_declared_symbol():
    _lib.build()
    lib = _lib.load()
    fns = _header_functions()
    assert len(fns) >= 25
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.ab_version() >= 100


def test_ctypes_table_mirrors_header():
    fns = _header_functions()
    assert set(fns) == set(_lib.SIGNATURES), set(fns) ^ set(_lib.SIGNATURES)
    for name, n in fns.items():
        assert len(_lib.SIGNATURES[name][1]) == n, (name, n, len(_lib.SIGNATURES[name][1]))


def test_pure_queries_work_without_gpu():
    r = _lib.query("ab_moe_max_rows", 4096, 2, 8, 640, 128)
    assert r % 128 == 0 and r >= 8 * 640
    import ctypes
    t, s, n, m = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(_lib.SCAN_SINGLE_PASS)
    ws = ctypes.c_size_t()
    plan = lambda *a: _lib.query("ab_selective_scan_plan", *a, ctypes.byref(m), ctypes.byref(t), ctypes.byref(s), ctypes.byref(n), ctypes.byref(ws))
    rc = plan(1, 65536, 512, _lib.AB_BF16)
    assert rc == 0 and t.value % 4 == 0 and 512 % s.value == 0 and n.value == -(-65536 // t.value)
    # pipelined schedule: per batch one saved state per run of 4 tokens, then delta [L, H] (rows of Di floats); a batch with more chains than scanner CTAs falls back
    m.value = _lib.SCAN_PIPELINED
    rc = plan(1, 65536, 512, _lib.AB_BF16)
    assert rc == 0 and m.value == _lib.SCAN_PIPELINED and s.value == 64 and n.value == 65536 // 4 + 65536 // 16 and ws.value > 0
    m.value = _lib.SCAN_PIPELINED
    rc = plan(64, 4096, 512, _lib.AB_BF16)
    assert rc == 0 and m.value == _lib.SCAN_SINGLE_PASS and n.value == -(-4096 // t.value)
    m.value = _lib.SCAN_SINGLE_PASS
    rc = plan(1, 16, 20, _lib.AB_BF16)
    assert rc != 0 and "tiling" in _lib.last_error()


def test_scan_plan_pipelined_host_logic():
    """Plan of the pipelined scan schedule (pure host code): which shapes it covers, the saved-state rows and that the
    workspace grows with the problem."""
    import ctypes

    def plan(B, L, Di, dtype, mode):
        t, s, n, m = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int(mode)
        ws = ctypes.c_size_t()
        rc = _lib.query("ab_selective_scan_plan", B, L, Di, dtype, ctypes.byref(m), ctypes.byref(t), ctypes.byref(s), ctypes.byref(n),
                        ctypes.byref(ws))
        return rc, m.value, t.value, s.value, n.value, ws.value

    P = _lib.SCAN_PIPELINED
    prev_ws = 0
    for L in (1, 5, 63, 64, 1000, 4096, 65536):
        rc, mode, T, Cs, n, ws = plan(1, L, 512, _lib.AB_BF16, P)
        assert rc == 0 and mode == P and Cs == 64 and T == 64
        Hp = 32
        assert n == -(-L // 4) + -(-L * Hp // 512)          # run states, then delta rows, in rows of Di floats
        assert ws >= prev_ws and ws > 0
        prev_ws = ws
    # fp32 activations: half as many tokens per tile; odd head counts: one slab over the whole width when it fits a CTA row
    rc, mode, T, Cs, n, ws = plan(2, 4096, 512, _lib.AB_F32, P)
    assert rc == 0 and mode == P and (T, Cs) == (32, 64)
    rc, mode, T, Cs, n, ws = plan(8, 4096, 176, _lib.AB_BF16, P)
    assert rc == 0 and mode == P and Cs == 176 and T == 20 and n == 1024 + -(-4096 * 12 // 176)
    # more chains than a quarter of the persistent grid, or a width whose slab would idle most lanes: another schedule
    for args in ((64, 4096, 512), (1, 4096, 16 * 17)):
        rc, mode, T, Cs, n, ws = plan(*args, _lib.AB_BF16, P)
        assert rc == 0 and mode == _lib.SCAN_SINGLE_PASS and n == -(-args[1] // T)
