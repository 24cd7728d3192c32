"""The C-ABI library: it loads without a GPU, exports every symbol include/scg_b200.h declares, the
ctypes binding covers exactly those symbols, and the product path refuses to run without CUDA (no
CPU fallback, no route through oracle/)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scg_b200.h")
PKG = os.path.join(ROOT, "skill-chaining-with-graphs_b200")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import skill_chaining_with_graphs_b200 as scg
    lib = ctypes.CDLL(scg.LIB_PATH)
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/scg_b200.h but not exported"
    assert sorted(scg.EXPORTS) == names, "ctypes binding and header disagree"


def test_version_error_strings_and_argument_checks_without_gpu():
    import skill_chaining_with_graphs_b200 as scg
    lib = scg.load_library()
    assert lib.scg_version() >= 100
    assert lib.scg_error_string(0) == b"ok"
    assert b"invalid" in lib.scg_error_string(-1) and b"limit" in lib.scg_error_string(-3)
    assert lib.scg_launch_count() == 0
    # argument validation happens before any CUDA call
    assert lib.scg_map_create(None, None, 0, 0.02, 0.9, 0.2, 0.04, None, 0, 0, None) == -1
    assert lib.scg_step(None, 4, *([None] * 11), 1, None) == -1
    assert lib.scg_ctx_create(9, 4, ctypes.byref(ctypes.c_void_p())) == -3
    assert lib.scg_ctx_create(3, 99, ctypes.byref(ctypes.c_void_p())) == -3
    assert lib.scg_agent_step(None, None, None, None) == -1


def test_packed_weight_slot_sizes():
    """The packed copy is [K][scg_packed_slot_floats(order)]: pairs of features x 12 floats, padded so that consecutive
    slots start 16 bytes apart modulo 128 (lanes on different slots then read different shared-memory banks)."""
    import skill_chaining_with_graphs_b200 as scg
    lib = scg.load_library()
    for order in range(1, 6):
        n1 = order + 1
        pairs = n1 ** 3 * ((n1 + 1) // 2)
        floats = lib.scg_packed_slot_floats(order)
        assert floats * 4 >= pairs * 48 and floats * 4 - pairs * 48 < 128
        assert (floats * 4) % 128 == 16 and floats % 4 == 0
    assert lib.scg_packed_slot_floats(0) == 0 and lib.scg_packed_slot_floats(6) == 0
    assert lib.scg_packed_slot_floats(3) == 1540 and lib.scg_packed_slot_floats(5) == 7780


def test_agent_struct_matches_header_field_order():
    from skill_chaining_with_graphs_b200._lib import AgentStruct
    src = open(HEADER).read()
    body = re.search(r"typedef struct scg_agent \{(.*?)\} scg_agent_t;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1] if not decl.startswith("const") else decl.split(None, 2)[2]
        fields += [n.strip().lstrip("*").strip() for n in names.split(",")]
    got = [("lam" if False else n) for n, _ in AgentStruct._fields_]
    assert [f.replace("lambda", "lam") for f in fields] == got


def test_no_cpu_fallback_and_no_oracle_in_product():
    import torch
    import skill_chaining_with_graphs_b200 as scg
    for fn in sorted(os.listdir(PKG)) + [os.path.join("csrc", f) for f in sorted(os.listdir(os.path.join(PKG, "csrc")))]:
        p = os.path.join(PKG, fn)
        if os.path.isfile(p) and p.endswith((".py", ".cu", ".cuh", ".h")):
            text = open(p).read()
            assert not re.search(r"^\s*(import|from)\s+oracle", text, flags=re.M), f"{fn} imports the oracle"
    if not torch.cuda.is_available():
        with pytest.raises(scg.ScgError):
            scg.PinballEnv("easy", 4)
        with pytest.raises(scg.ScgError):
            scg.OptionSet(1, 3, 4)
        with pytest.raises(scg.ScgError):
            scg.SkillChainAgent(scg.AgentConfig())


def test_env_slices_cover_the_batch():
    from skill_chaining_with_graphs_b200.sync import env_slice
    for total, world in ((65536, 8), (10, 3), (7, 8), (1 << 20, 8)):
        sl = [env_slice(total, r, world) for r in range(world)]
        assert sl[0][0] == 0 and sl[-1][1] == total
        assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in sl]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        env_slice(8, 8, 8)
