"""CPU check for names that only fail on the GPU box: every global name a function of bench.py, __graft_entry__.py, the
package or the profile tools loads must exist at module level or in builtins (the GPU arms cannot be run here, so a
missing import inside them would otherwise surface at round end)."""
import builtins
import dis
import glob
import os
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ([os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
         + sorted(glob.glob(os.path.join(ROOT, "gromacs_b200", "*.py")))
         + sorted(glob.glob(os.path.join(ROOT, "profiles", "tools", "*.py"))))


def code_objects(code):
    yield code
    for c in code.co_consts:
        if isinstance(c, types.CodeType):
            yield from code_objects(c)


@pytest.mark.parametrize("path", FILES, ids=[os.path.relpath(f, ROOT) for f in FILES])
def test_global_names_resolve(path):
    src = open(path).read()
    top = compile(src, path, "exec")
    module_names = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__annotations__", "__module__", "__qualname__"}
    for code in code_objects(top):
        for ins in dis.get_instructions(code):
            if ins.opname in ("STORE_NAME", "STORE_GLOBAL", "IMPORT_NAME") and code is top:
                module_names.add(ins.argval.split(".")[0])
            if ins.opname in ("STORE_NAME", "STORE_GLOBAL"):
                module_names.add(ins.argval)
    missing = []
    for code in code_objects(top):
        if code is top:
            continue
        for ins in dis.get_instructions(code):
            # class bodies see the names they define themselves
            if ins.opname == "LOAD_NAME" and any(i.opname == "STORE_NAME" and i.argval == ins.argval for i in dis.get_instructions(code)):
                continue
            if ins.opname in ("LOAD_GLOBAL", "LOAD_NAME") and ins.argval not in module_names:
                missing.append("%s:%s uses undefined global %r" % (code.co_name, ins.positions.lineno, ins.argval))
    assert not missing, "\n".join(missing)
