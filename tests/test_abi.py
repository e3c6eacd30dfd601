"""CPU tests of the drop-in boundary: the shared library loads without a GPU, exports every symbol
include/quda.h + include/quda_b200_ext.h declare, and its parameter structs are byte-identical to
the reference's (compiled side by side when /root/reference is present) and to the ctypes mirror."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q  # noqa: E402

REF_INC = "/root/reference/include"


def _declared_functions():
    names = []
    for hdr in ("quda.h", "quda_b200_ext.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"^\s*(?:[A-Za-z_][\w\s\*]*?)\b(\w+(?:Quda|QudaB200)\w*)\s*\(", src, flags=re.M)
    return sorted(set(n for n in names if not n.startswith("Quda")))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(q.LIB_PATH), "libquda_b200.so not built: run __graft_entry__.build()"
    L = C.CDLL(q.LIB_PATH)
    declared = _declared_functions()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    for name in q.EXPORTS:
        assert name in declared, f"{name} bound in api.py but not declared in include/"


def _layout_program(include_dir, header):
    fields = {
        "QudaGaugeParam": [f[0] for f in q.QudaGaugeParam._fields_],
        "QudaInvertParam": [f[0] for f in q.QudaInvertParam._fields_],
        "QudaMultigridParam": [f[0] for f in q.QudaMultigridParam._fields_],
    }
    body = ['#include <stdio.h>', '#include <stddef.h>', f'#include <{header}>', 'int main(){']
    for s, fl in fields.items():
        body.append(f'printf("{s} %zu\\n", sizeof({s}));')
        for f in fl:
            body.append(f'printf("{s}.{f} %zu\\n", offsetof({s}, {f}));')
    body.append('return 0;}')
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "layout.c")
        open(src, "w").write("\n".join(body))
        exe = os.path.join(td, "layout")
        subprocess.check_call(["gcc", "-I", include_dir, "-o", exe, src])
        out = subprocess.check_output([exe]).decode()
    return dict(line.split() for line in out.strip().splitlines())


def _ctypes_layout():
    d = {}
    for cls in (q.QudaGaugeParam, q.QudaInvertParam, q.QudaMultigridParam):
        d[cls.__name__] = str(C.sizeof(cls))
        for name, _ in cls._fields_:
            d[f"{cls.__name__}.{name}"] = str(getattr(cls, name).offset)
    return d


def test_struct_layout_header_vs_ctypes():
    ours = _layout_program(os.path.join(ROOT, "include"), "quda.h")
    assert ours == _ctypes_layout()


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference tree not present (GPU box)")
def test_struct_layout_identical_to_reference():
    ours = _layout_program(os.path.join(ROOT, "include"), "quda.h")
    ref = _layout_program(REF_INC, "quda.h")
    assert ours == ref


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="reference tree not present (GPU box)")
def test_enum_values_identical_to_reference():
    names = sorted(set(re.findall(r"\b(QUDA_[A-Z0-9_]+)\b", open(os.path.join(ROOT, "include", "quda_b200_enums.h")).read())))
    names = [n for n in names if n not in ("QUDA_B200_ENUMS_H", "QUDA_INVALID_ENUM")]
    prog = ['#include <stdio.h>', '#include <HDR>', 'int main(){'] + [f'printf("{n} %d\\n", (int){n});' for n in names] + ['return 0;}']

    def run(inc, hdr):
        with tempfile.TemporaryDirectory() as td:
            src = os.path.join(td, "e.c")
            open(src, "w").write("\n".join(prog).replace("HDR", hdr))
            exe = os.path.join(td, "e")
            subprocess.check_call(["gcc", "-I", inc, "-o", exe, src])
            return subprocess.check_output([exe]).decode()

    assert run(os.path.join(ROOT, "include"), "quda_b200_enums.h") == run(REF_INC, "enum_quda.h")
    # and the Python constants
    vals = dict(l.split() for l in run(os.path.join(ROOT, "include"), "quda_b200_enums.h").strip().splitlines())
    for n, v in vals.items():
        if hasattr(q, n):
            assert getattr(q, n) == int(v), n


def test_param_constructors_poison_and_defaults():
    L = q.lib()  # loading needs no GPU; no compute call is made here
    g = L.newQudaGaugeParam()
    assert g.X[0] == q.QUDA_INVALID_ENUM and g.reconstruct == q.QUDA_INVALID_ENUM and g.location == q.QUDA_CPU_FIELD_LOCATION
    p = L.newQudaInvertParam()
    assert p.dslash_type == q.QUDA_INVALID_ENUM and p.num_src == 1 and p.omega == 1.0 and p.max_res_increase == 1
    assert p.preconditioner is None and p.precondition_cycle == 1
    m = L.newQudaMultigridParam()
    assert m.n_level == q.QUDA_INVALID_ENUM and m.global_reduction[0] == q.QUDA_BOOLEAN_YES


def test_error_model_is_process_exit():
    """Errors end the process with a message (reference: errorQuda -> comm_abort), no return code."""
    code = ("import sys; sys.path.insert(0, %r); import quda_b200 as q; L=q.lib(); "
            "p=L.newQudaInvertParam(); L.dslashQuda(None, None, p, 0)") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 1
    assert "ERROR: QUDA not initialized" in (r.stderr + r.stdout)
