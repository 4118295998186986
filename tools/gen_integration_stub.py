"""Prints the ctypes mirror of `struct dk_gemm_args` for INTEGRATION.md §2, generated from diffusionkit_b200/_lib.py
(which tests/test_host_cpu.py checks field by field against include/dkb200.h) so that the document cannot drift."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BEGIN, END = "<!-- BEGIN GENERATED dk_gemm_args (tools/gen_integration_stub.py) -->", "<!-- END GENERATED dk_gemm_args -->"
NAMES = {C.c_int: "C.c_int", C.c_void_p: "C.c_void_p", C.c_longlong: "C.c_longlong", C.c_float: "C.c_float"}


def block() -> str:
    from diffusionkit_b200._lib import GemmArgs

    rows, line = [], "    _fields_ = ["
    for name, typ in GemmArgs._fields_:
        item = f'("{name}", {NAMES[typ]}), '
        if len(line) + len(item) > 110:
            rows.append(line.rstrip())
            line = "                " + item
        else:
            line += item
    rows.append(line.rstrip().rstrip(",") + "]")
    body = "\n".join(rows)
    return (f"{BEGIN}\n```python\nclass GemmArgs(C.Structure):            # mirrors struct dk_gemm_args, field for field\n"
            f"{body}\n```\n{END}")


if __name__ == "__main__":
    print(block())
