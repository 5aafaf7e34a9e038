"""CPU: the C-ABI shared library loads and exports exactly what include/m3vit_moe.h
declares, with the arities the ctypes binding assumes (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "m3vit_moe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(m3_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def built_lib():
    from m3vit_b200 import build
    return build.build()


def test_header_declares_functions():
    fns = header_functions()
    assert len(fns) >= 25 and "m3_gate_fwd" in fns and "m3_ffn_fwd" in fns


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in m3vit_moe.h but not exported"


def test_ctypes_signatures_match_header(built_lib):
    from m3vit_b200 import _lib
    fns = header_functions()
    assert set(fns) == set(_lib.SIGNATURES), set(fns) ^ set(_lib.SIGNATURES)
    for name, n in fns.items():
        assert len(_lib.SIGNATURES[name][1]) == n, f"{name}: header has {n} args, binding {len(_lib.SIGNATURES[name][1])}"
    lib = _lib.load()
    assert lib.m3_abi_version() == 2
    assert lib.m3_status_string(-2).decode() == "unsupported shape"


def test_pure_host_entry_points(built_lib):
    """Entry points that do no device work can be called without a GPU."""
    from m3vit_b200 import _lib
    lib = _lib.load()
    assert lib.m3_route_max_rows(2402, 4, 16, 128) == (2402 * 4 + 16 * 127 + 127) // 128 * 128
    assert lib.m3_route_max_tiles(2402, 4, 16, 128) * 128 == lib.m3_route_max_rows(2402, 4, 16, 128)
    assert lib.m3_gate_num_partials(2402, 16) == (2402 + 15) // 16
    assert lib.m3_gate_num_partials(2402, 12) < 0           # unsupported expert count
    assert lib.m3_route_plan_workspace_bytes(2402, 4, 16) == ((2402 * 4 + 2047) // 2048) * 16 * 4
    assert lib.m3_ffn_workspace_bytes(0, 1024, 64, 128, 16, 0) == 1024 * 128 * 4


def test_null_arguments_are_rejected_not_dereferenced(built_lib):
    from m3vit_b200 import _lib
    lib = _lib.load()
    assert lib.m3_dispatch_fwd(None, 0, None, None, None, 1, 1, 64, 16, None, 0, None) == -1
    assert lib.m3_ffn_fwd(0, None, None, None, 128, 16, 64, 64, None, None, None, None, None, None, None, 0, None) == -1


def test_expert_parallel_entry_points_validate_before_touching_the_device(built_lib):
    """The fused-return entry points (round 2) reject bad argument sets on the host: no launch, no dereference."""
    from m3vit_b200 import _lib
    lib = _lib.load()
    P = 0x1000      # a 16-byte aligned fake device address (never dereferenced on these paths)
    # return store needs BOTH the row origins and the peers' return buffers
    assert lib.m3_ep_ffn_fwd(1, P, P, P, 256, 2, 64, 64, P, P, P, P, None, None, P, P, 1 << 20, 0.0, None, None) == -1
    assert lib.m3_ep_ffn_fwd(1, P, P, P, 256, 2, 64, 64, P, P, P, P, None, P, None, P, 1 << 20, 0.0, None, None) == -1
    # ... and lives in the tcgen05 epilogue only: fp32 queues are refused, not silently run without the return
    assert lib.m3_ep_ffn_fwd(0, P, P, P, 256, 2, 64, 64, P, P, P, P, None, P, P, P, 1 << 20, 0.0, None, None) == -4
    assert lib.m3_ep_ffn_bwd(0, P, P, P, P, P, P, 256, 2, 64, 64, P, P, P, P, P, P, P, P, P, P, P, 1 << 20, 0.0, None, 3,
                             None) == -4
    # parts of the split backward: 1 = data gradients, 2 = weight gradients, 3 = both
    assert lib.m3_ffn_bwd_parts(1, P, P, P, P, P, P, 256, 2, 64, 64, P, P, P, P, P, P, P, P, P, P, 1 << 20, 0.0, None, 0,
                                None) == -1
    # row origins and the sources' inverse plans come together; slots must fit 24 bits
    assert lib.m3_ep_plan(P, P, P, 0, 2, 2, 16, 2, 256, 512, P, P, P, P, P, None, None, None, P, None) == -1
    assert lib.m3_ep_plan(P, P, P, 0, 2, 2, 1 << 23, 4, 256, 512, P, P, P, P, P, None, None, P, P, None) == -1
    # an empty call is a no-op
    assert lib.m3_ep_ffn_fwd(1, P, P, P, 0, 2, 64, 64, P, P, P, P, None, P, P, P, 1 << 20, 0.0, None, None) == 0
