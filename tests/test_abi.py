"""The C-ABI library builds, loads and exports every symbol declared in include/tce_b200.h (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tce_b200.h")


@pytest.fixture(scope="module")
def lib():
    from tce_rl_b200 import _build, _lib
    _build.build()
    return _lib.load()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tce_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in tce_b200.h but not exported"


def test_binding_covers_header(lib):
    from tce_rl_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_binding_argument_counts_match_header():
    """Every ctypes signature has exactly as many arguments as the declaration in the header (a short argtypes list
    makes ctypes silently pass garbage)."""
    from tce_rl_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    seen = 0
    for m in re.finditer(r"\b(?:int|size_t|void|const char \*)\s*\**(tce_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert len(_lib.SIGNATURES[name][1]) == n, name
        seen += 1
    assert seen == len(_lib.SIGNATURES)


def test_version_and_errors(lib):
    assert lib.tce_version() >= 100
    assert lib.tce_strerror(0) == b"ok"
    assert b"unsupported" in lib.tce_strerror(-2)
    # argument validation happens before any CUDA call
    assert lib.tce_gae(None, None, None, None, 1.0, 0.95, 1, None, None, 4, 10, None) == -1
    assert lib.tce_prodmp_tables_create(None, None, None) == -1


def test_mp_cfg_layout():
    from tce_rl_b200._lib import MpCfg
    assert ctypes.sizeof(MpCfg) == 8 * 4 + 8 * 8


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "tce_rl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "from .. import oracle" not in src and "/root/reference" not in src


def test_library_contains_the_sm100_instructions_the_design_names():
    """DESIGN.md section 4 / profiles/r02_sass_evidence.txt: the factor-streaming kernels request their data with bulk
    asynchronous copies (UBLKCP + mbarrier SYNCS), the common-grid trajectory kernel uses packed fp32x2 FMAs (FFMA2), and
    nothing is on the tensor cores.  Checked on the object files of the in-tree build (cuobjdump, no GPU needed)."""
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    build = os.path.join(ROOT, "tce_rl_b200", "build")
    objs = {f: os.path.join(build, f) for f in ("tce_gauss.o", "tce_traj.o")}
    if not os.path.exists(tool) or not all(os.path.exists(p) for p in objs.values()):
        pytest.skip("cuobjdump or the object files of the in-tree build are not here")
    sass = {f: subprocess.run([tool, "-sass", p], capture_output=True, text=True, check=True).stdout for f, p in objs.items()}
    assert "UBLKCP.S.G" in sass["tce_gauss.o"] and "SYNCS.ARRIVE.TRANS64" in sass["tce_gauss.o"]
    assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in sass["tce_gauss.o"]
    assert "FFMA2" in sass["tce_traj.o"]
    opcodes = {op for s in sass.values() for op in re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", s)}
    assert "FFMA" in opcodes and not [op for op in opcodes if "MMA" in op]


def test_traffic_json_entries_name_existing_sources():
    """bench.py reports ``roofline.traffic`` from profiles/r02_ncu_traffic.json only while the sources of THAT kernel
    are unchanged: every prefix of the source map points at a file that exists, and every entry carries its hash."""
    import json
    import bench
    csrc = os.path.join(ROOT, "tce_rl_b200", "csrc")
    assert all(os.path.exists(os.path.join(csrc, f)) for _, f in bench._SRC_OF)
    tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
    for name, grids in tj["kernels"].items():
        assert any(name.startswith(p) for p, _ in bench._SRC_OF), name
        assert all(len(e["src_hash"]) == 16 and e["dram_bytes"] > 0 for e in grids.values())
    assert bench.src_hash_of("tce_proj_kl_entropy_fwd_sigma_vec") != bench.src_hash_of("tce_prodmp_traj_fwd")


def test_header_is_plain_c_and_links_against_the_library(lib, tmp_path):
    """include/tce_b200.h is what a cgo / JNI / ctypes binding reads: it must compile as C99 and as C++, and a C program
    that only knows the header must link against libtce_b200.so and get answers from the calls that need no GPU."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", HEADER], check=True)
    gxx = shutil.which("g++")
    if gxx:
        subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-x", "c++", HEADER], check=True)
    src = tmp_path / "demo.c"
    src.write_text('#include <stdio.h>\n#include "tce_b200.h"\n'
                   'int main(void) {\n'
                   '  /* NULL pointers: rejected before any CUDA call */\n'
                   '  int rc = tce_mvn_rsample(0, 0, 0, 0, 0, 0, 0, 4, 3, 0);\n'
                   '  printf("%d %d %s\\n", tce_version(), rc, tce_strerror(rc));\n'
                   '  return rc == TCE_ERR_INVALID_ARGUMENT ? 0 : 1;\n}\n')
    libdir = os.path.join(ROOT, "tce_rl_b200")
    exe = tmp_path / "demo"
    subprocess.run([gcc, "-std=c99", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe), "-L", libdir,
                    "-l:libtce_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert int(out.stdout.split()[0]) >= 100
