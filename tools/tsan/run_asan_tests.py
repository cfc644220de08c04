import sys, ctypes as C, pathlib, tempfile, inspect
sys.path.insert(0, '/root/repo')
import numpy as np
from scasa_b200 import lib as _l
L = C.CDLL('/tmp/scasa_tsan/libhost_asan.so')
L.scasa_last_error.restype = C.c_char_p
import re
src = open('/root/repo/scasa_b200/lib.py').read()
src = re.sub(r",\n\s+", ", ", src)                      # argtypes lists that continue on the next line
# reuse the argtypes lib.py declares for the host-only entry points
vp, i32p, i64p, f64p, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
for m in re.finditer(r"^    L\.(scasa_(?:bfh|bus|write|format|csr_transpose)[a-z0-9_]*)\.(argtypes|restype) = (.*)$", src, re.M):
    setattr(getattr(L, m.group(1)), m.group(2), eval(m.group(3)))
_l.load = lambda: L
_l._lib = L
import tests.test_bfh as tb
import tests.test_writers as tw
n = 0
for mod in (tb, tw):
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith('test_'): continue
        sig = inspect.signature(fn)
        marks = getattr(fn, 'pytestmark', [])
        combos = [{}]
        for m in marks:
            if m.name == 'parametrize':
                keys = [k.strip() for k in m.args[0].split(',')]
                new = []
                for c in combos:
                    for v in m.args[1]:
                        vv = v if isinstance(v, (tuple, list)) and len(keys) > 1 else (v,)
                        d = dict(c); d.update(dict(zip(keys, vv))); new.append(d)
                combos = new
        for c in combos:
            kw = dict(c)
            if 'tmp_path' in sig.parameters: kw['tmp_path'] = pathlib.Path(tempfile.mkdtemp())
            try:
                fn(**kw); n += 1
            except TypeError as e:
                print('skip', name, e)
print('ran', n, 'tests under ASan/UBSan')
