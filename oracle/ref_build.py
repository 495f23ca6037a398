"""Make the reference's OWN sources executable in this container -- TEST INFRASTRUCTURE ONLY.

/root/reference/{wavenet,faster_wavenet,data}.py are Python 2 over Chainer 2; neither exists here.  This script applies
a MECHANICAL py2 -> py3 transform (listed below, nothing else is touched) and writes the results to oracle/_ref/
(git-ignored: derived copies of the reference never enter the history).  Together with the NumPy `chainer` stand-in in
oracle/chainer_shim/ the reference's control flow then runs for real; tests/golden/make_ref_golden.py uses it to produce
the committed fixtures tests/golden/ref_*.npz that pin oracle/wavenet_oracle.py and the CUDA path.

Transform rules
  1. `print x, y`            -> `print(x, y)`                      (wavenet.py:163,165,630,636)
  2. `xrange(`               -> `range(`                           (wavenet.py:285,289,412,414; data.py:27,30)
  3. `.iteritems()`          -> `.items()`                         (wavenet.py:152,158,164,169)
  4. `padded_x_width / self.dilation` -> `//`                      (wavenet.py:314: py2 integer division of two ints)
  5. data.py only: CRLF -> LF; a leading " \\t" -> "\\t"          (data.py:7-9: py2 expands the tab to column 8, so
     space+tab IS one indentation level; py3 rejects the mix as TabError)
  6. data.py only: `signal /= max` -> `signal = _py2_idiv(signal, max)`, where _py2_idiv is floor division for integer
     arrays and true division for float arrays -- what `ndarray.__idiv__` did under Python 2 (data.py:17, quirk Q5)

Runs only where /root/reference exists (this container); the GPU box uses the committed fixtures.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WN_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")

PY2_IDIV = '''
def _py2_idiv(a, b):
	# ndarray.__idiv__ under Python 2: classic division in place -- floor division for integer arrays
	import numpy as _np
	if _np.issubdtype(a.dtype, _np.integer):
		_np.floor_divide(a, _np.int64(b), out=a, casting="unsafe")
	else:
		a /= b
	return a
'''


def transform(name, src):
    out = []
    if name == "data.py":
        src = src.replace("\r\n", "\n")
    for line in src.split("\n"):
        if name == "data.py" and line.startswith(" \t"):
            line = line[1:]
        m = re.match(r"^(\s*)print (.+)$", line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = line.replace("xrange(", "range(").replace(".iteritems()", ".items()")
        line = line.replace("padded_x_width / self.dilation", "padded_x_width // self.dilation")
        if name == "data.py" and line.strip() == "signal /= max":
            line = line.replace("signal /= max", "signal = _py2_idiv(signal, max)")
        out.append(line)
    text = "\n".join(out)
    if name == "data.py":
        text = text.replace("import numpy as np\n", "import numpy as np\n" + PY2_IDIV, 1)
    return text


def build(verbose=True):
    if not os.path.isdir(REF):
        raise RuntimeError("reference sources not found at %s (ref_build only runs in the build container)" % REF)
    os.makedirs(OUT, exist_ok=True)
    for name in ("wavenet.py", "faster_wavenet.py", "data.py"):
        with open(os.path.join(REF, name), "r", newline="") as f:
            src = f.read()
        dst = os.path.join(OUT, name)
        with open(dst, "w") as f:
            f.write(transform(name, src))
        compile(open(dst).read(), dst, "exec")    # must at least parse under Python 3
        if verbose:
            print("ref_build: %s -> %s" % (os.path.join(REF, name), dst))
    return OUT


def import_reference():
    """Returns the transformed reference modules (wavenet, faster_wavenet, data) running on the chainer stand-in.
    Rebuilds oracle/_ref/ when /root/reference is present; otherwise (GPU box) uses the files that travelled with the tree."""
    if os.path.isdir(REF):
        out = build(verbose=False)
    elif all(os.path.isfile(os.path.join(OUT, n)) for n in ("wavenet.py", "faster_wavenet.py", "data.py")):
        out = OUT
    else:
        raise RuntimeError("oracle/_ref/ is empty and %s does not exist: run oracle/ref_build.py in the build container" % REF)
    shim = os.path.join(HERE, "chainer_shim")
    for p in (shim, out):
        if p not in sys.path:
            sys.path.insert(0, p)
    for mod in ("wavenet", "faster_wavenet", "data"):
        sys.modules.pop(mod, None)
    import wavenet as R
    import faster_wavenet as RF
    import data as RD
    return R, RF, RD


if __name__ == "__main__":
    build()
