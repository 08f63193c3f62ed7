"""CPU: the C-ABI shared library loads and exports every symbol include/argus_b200.h declares (no compute calls),
and the product package never imports the oracle."""
import ast
from pathlib import Path

from argus_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.argus_version() >= 100


def test_errors_do_not_cross_the_boundary_as_exceptions():
    lib = _lib.load()
    # no GPU / wrong arguments: a non-zero status plus a message, never a crash
    status = lib.argus_model_tensor_info(None, 0, 0, None, 0, None, None, None, None)
    assert status != 0
    assert b"null model" in lib.argus_last_error_string()


def test_product_never_imports_the_oracle():
    for path in (ROOT / "argus_b200").rglob("*.py"):
        tree = ast.parse(path.read_text())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.module:
                names = [node.module]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path
    for path in (ROOT / "argus_b200" / "csrc").glob("*"):
        assert "oracle/" not in path.read_text().replace("oracle/se3_loss.py", "").replace("oracle/augment.py", "").replace("oracle/pil_arc.py", ""), path


def test_every_kernel_waits_on_its_programmatic_dependency():
    """Kernels are launched with a programmatic-dependent-launch edge (runtime.h::launch_kernel): a kernel that does not
    start with pdl_prologue() (griddepcontrol.wait) would run before its producer has finished. Also: no raw <<<>>>
    launch is left that would bypass the launch accounting."""
    import re

    for path in sorted((ROOT / "argus_b200" / "csrc").glob("*.cu*")):
        text = re.sub(r"//[^\n]*", "", path.read_text())   # comments may contain parentheses / braces
        assert "<<<" not in text, f"{path.name}: raw kernel launch"
        for m in re.finditer(r"__global__", text):
            brace = text.index("{", m.end())
            semi = text.find(";", m.end())
            if 0 <= semi < brace:
                continue   # declaration only
            body = text[brace + 1:brace + 200].lstrip()
            if body.startswith("pdl_begin("):
                # deferred form (ptx.cuh): pdl_begin(0) waits at once; pdl_begin(1) triggers and the wait stands after the
                # data-independent prologue, before the role dispatch -- both take the same flag
                end = text.find("__global__", m.end())
                span = text[brace:end if end > 0 else len(text)]
                flag = re.match(r"pdl_begin\(([^)]*)\)", body).group(1)
                assert f"pdl_wait_deferred({flag});" in span, f"{path.name}: pdl_begin() without pdl_wait_deferred()"
                continue
            assert body.startswith("pdl_prologue();"), f"{path.name}: kernel at offset {m.start()} lacks pdl_prologue()"
