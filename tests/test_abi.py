"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/kpgnn.h declares.
No compute calls are made (there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(kp_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol(lib):
    from kpgnn_b200 import _lib
    declared = _declared()
    assert len(declared) >= 8
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), "missing export " + name
    # and the Python binding covers them all
    assert declared == set(_lib._SIGNATURES.keys())


def test_abi_version_and_error_string(lib):
    from kpgnn_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "kpgnn.h")).read()
    declared = int(re.search(r"#define\s+KPGNN_ABI_VERSION\s+(\d+)", hdr).group(1))
    assert lib.kp_abi_version() == _lib.ABI_VERSION == declared
    assert isinstance(lib.kp_last_error(), bytes)
    assert lib.kp_launch_count() >= 0


def test_argument_validation_without_gpu(lib):
    """Entry points reject bad descriptors before touching the device."""
    import ctypes as C
    from kpgnn_b200 import _lib
    n = C.c_size_t(0)
    assert lib.kp_plan_workspace_bytes(10, 20, 0, C.byref(n)) != 0
    assert b"bad arguments" in lib.kp_last_error()
    assert lib.kp_plan_workspace_bytes(10, 20, 3, C.byref(n)) == 0 and n.value > 0
    d = _lib.AggDesc()
    assert lib.kp_agg_forward(C.byref(d), None, None) != 0


def test_product_never_imports_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kpgnn_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_struct_layouts_match_header(tmp_path):
    """Every ctypes mirror in kpgnn_b200/_lib.py has the size and the field offsets the C compiler gives the struct
    declared in include/kpgnn.h (a silent drift here would make the kernels read garbage descriptors)."""
    import subprocess
    from kpgnn_b200 import _lib
    mirrors = {"kp_plan_input": _lib.PlanInput, "kp_agg_desc": _lib.AggDesc, "kp_extract_input": _lib.ExtractInput,
               "kp_tsum_desc": _lib.TsumDesc, "kp_dense_desc": _lib.DenseDesc, "kp_theta_batch": _lib.ThetaBatch,
               "kp_pgrad_desc": _lib.PgradDesc, "kp_attn_desc": _lib.AttnDesc, "kp_wire_desc": _lib.WireDesc, "kp_peer_desc": _lib.PeerDesc, "kp_head_desc": _lib.HeadDesc, "kp_fold_desc": _lib.FoldDesc,
               "kp_fold_grads": _lib.FoldGrads}
    header = open(os.path.join(ROOT, "include", "kpgnn.h")).read()
    declared = set(re.findall(r"^\}\s*(kp_[a-z0-9_]+)\s*;", header, flags=re.M))
    # structs that never cross the Python boundary as ctypes objects are packed with numpy/torch instead
    assert declared - set(mirrors) <= {"kp_adam_tensor"}, declared - set(mirrors)
    lines = ['#include <stdio.h>', '#include "kpgnn.h"', "int main(void) {"]
    for cname, cls in mirrors.items():
        lines.append('  printf("%s.sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    c_layout = dict(line.split() for line in out.splitlines())
    for cname, cls in mirrors.items():
        assert int(c_layout[cname + ".sizeof"]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(c_layout["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)
