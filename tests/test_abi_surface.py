"""The C-ABI library builds, loads, and exports every symbol include/pyrite_b200.h declares (no compute)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "pyrite_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pyr_[a-z_0-9]+)\s*\(", text)) - {"pyr_progress_cb"})


def test_library_exports_every_declared_symbol():
    from pyrite_b200 import api

    lib = api.load_library()
    names = declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/pyrite_b200.h but not exported"
    assert sorted(api.EXPORTS) == names, "pyrite_b200.api.EXPORTS and the header disagree"


def test_struct_layouts_match_the_header():
    from pyrite_b200 import api

    assert api.RAY_DTYPE.itemsize == 32 and api.HIT_DTYPE.itemsize == 20
    assert ctypes.sizeof(api.ProjectInfo) == 16 * 4
    assert ctypes.sizeof(api.RenderParams) == 40
    assert ctypes.sizeof(api.Counters) == 8 * 8 + 3 * 8 + 6 * 8


def test_version_string_names_the_target():
    from pyrite_b200 import api

    assert b"sm_100a" in api.load_library().pyr_version()


def test_init_fails_loudly_without_a_gpu():
    import torch

    from pyrite_b200 import api

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.PyriteError) as e:
        api.Renderer(0)
    assert e.value.status == 2 and "no CPU path" in str(e.value)


def test_missing_library_is_an_import_error(tmp_path):
    from pyrite_b200 import api

    with pytest.raises(ImportError):
        api.load_library(tmp_path / "nope.so")


def test_product_never_imports_the_oracle():
    pkg = ROOT / "pyrite_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.hpp")):
        text = p.read_text()
        assert "oracle_lib" not in text and "liboracle" not in text and "pyro_" not in text, f"{p} references the oracle"


def _build_abi_smoke():
    import subprocess

    build = ROOT / "tests" / "_build"
    build.mkdir(exist_ok=True)
    exe = build / "abi_smoke"
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), "-o", str(exe), str(ROOT / "tests" / "abi_smoke.c"), "-ldl", "-lm"],
                   check=True, capture_output=True)
    return exe


def test_header_compiles_as_plain_c_and_the_c_host_fails_loudly_without_a_gpu():
    """include/pyrite_b200.h is consumed by a C11 compiler (tests/abi_smoke.c, no Python in between); without a CUDA device
    the C host gets PYR_ERR_CUDA and a message from pyr_init - there is no CPU path to fall back to."""
    import subprocess

    import torch
    from conftest import scene_ir
    from oracle_lib import build as build_oracle

    exe = _build_abi_smoke()
    blob = ROOT / "tests" / "golden" / "cornell_64.ir"
    assert blob.read_bytes() == scene_ir("cornell"), "tests/golden/cornell_64.ir is stale (regenerate it from scene_ir('cornell'))"
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the GPU test runs the C host to completion")
    res = subprocess.run([str(exe), str(ROOT / "pyrite_b200" / "libpyrite_b200.so"), str(build_oracle()), str(blob)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 2 and "no CPU path" in res.stderr, (res.returncode, res.stderr)


@pytest.mark.gpu
def test_plain_c_host_runs_the_path_and_matches_the_oracle():
    """tests/abi_smoke.c: pyr_init -> pyr_project_load -> pyr_trace -> pyr_render -> pyr_film_develop from C, checked against the oracle."""
    import subprocess

    from oracle_lib import build as build_oracle

    exe = _build_abi_smoke()
    res = subprocess.run([str(exe), str(ROOT / "pyrite_b200" / "libpyrite_b200.so"), str(build_oracle()), str(ROOT / "tests" / "golden" / "cornell_64.ir")],
                         capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and "ABI SMOKE OK" in res.stdout, (res.stdout, res.stderr)


def test_comm_id_needs_no_gpu_and_reduce_needs_a_communicator():
    from pyrite_b200 import api

    a, b = api.Renderer.comm_unique_id(), api.Renderer.comm_unique_id()
    assert len(a) == 128 and a != b
