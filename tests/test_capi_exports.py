"""The C-ABI library: it loads, exports every symbol include/ismcts_b200.h declares, and its host game
helpers (no device needed) agree with the oracle.  No search is run here (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from ismcsolver_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ismcts_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ismcts_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(engine_lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(engine_lib, n), n
    assert sorted(capi.EXPORTS) == names


def test_struct_layouts_match_header(engine_lib):
    assert engine_lib.ismcts_state_size(capi.GAME_KW) == C.sizeof(capi.KwState) == 96
    assert engine_lib.ismcts_state_size(capi.GAME_MNK) == C.sizeof(capi.MnkState) == 72
    assert engine_lib.ismcts_state_size(9) == 0
    assert engine_lib.ismcts_move_stride(capi.GAME_KW) == 64
    assert engine_lib.ismcts_move_stride(capi.GAME_MNK) == 256
    assert C.sizeof(capi.SearchDesc) == 72 and C.sizeof(capi.SearchOut) == 56


def test_no_device_is_a_loud_error(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.EngineError) as e:
        capi.Engine(0)
    assert e.value.code == capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_host_helpers_match_oracle_over_full_games(engine_lib):
    rnd = np.random.default_rng(1)
    for stream in range(60):
        players = 2 + stream % 6
        a, b = O.kw_deal(77, stream, players), capi.kw_deal(77, stream, players)
        assert bytes(a) == bytes(b)
        ca, cb = [0], [0]
        for _ in range(400):
            la = O.kw_legal(a)
            mb = capi.kw_legal_mask(b)
            assert la == [c for c in range(52) if mb >> c & 1]
            if not la:
                break
            mv = int(rnd.choice(la))
            assert O.kw_apply(a, mv, 77, stream, ca) == capi.kw_apply(b, mv, 77, stream, cb) == 0
            assert bytes(a) == bytes(b) and ca == cb


def test_mnk_helpers_match_oracle(engine_lib):
    rnd = np.random.default_rng(2)
    for m, n, k in [(3, 3, 3), (5, 4, 4), (15, 15, 5), (16, 16, 6)]:
        for _ in range(10):
            a, b = O.mnk_init(m, n, k), capi.mnk_init(m, n, k)
            while True:
                la = O.mnk_legal(a)
                assert la == capi.mnk_legal(b)
                if not la:
                    break
                mv = int(rnd.choice(la))
                assert O.lib().orc_mnk_apply(C.byref(a), mv) == capi.mnk_apply(b, mv) == 0
                assert bytes(a) == bytes(b)


def _build_abi_example(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_example")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"), "-o", exe,
                    os.path.join(root, "tests", "c", "abi_example.c"), "-L" + os.path.join(root, "ismcsolver_b200"),
                    "-lismcts_b200", "-Wl,-rpath," + os.path.join(root, "ismcsolver_b200")], check=True)
    return exe


def test_header_is_plain_c_and_links(engine_lib, tmp_path):
    """include/ismcts_b200.h compiles as C11 with -Werror and a C program links against the library; without a device
    it is told so (ISMCTS_ERR_NO_DEVICE) instead of silently computing on the CPU."""
    import subprocess
    exe = _build_abi_example(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "no device" in out.stdout or "kernels launched" in out.stdout


@pytest.mark.gpu
def test_c_program_searches_on_the_gpu(engine_lib, tmp_path):
    import subprocess
    out = subprocess.run([_build_abi_example(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("position ") == 4 and "kernels launched" in out.stdout
