"""Test infrastructure only -- NOT part of the product path.

Builds a *temporary*, Python-3-importable copy of the handful of reference
modules on the FASST/SIMM hot path under a scratch directory (default
/tmp/pyfasst_ref_py3) so that the unmodified algorithm of the reference can be
*executed* in this container to produce golden vectors (oracle/make_golden.py).

Nothing produced here is ever committed: the reference sources stay where they
lie (/root/reference), the patched copies live in /tmp.  The patches are purely
syntactic / numpy-API-strictness fixes needed because the reference is
Python-2 + old-numpy code; each one is listed in PATCHES below
so a reviewer can check that none changes the arithmetic:

  * `print` statements (py2 syntax)           -> `pass` (output only)
  * implicit relative imports                 -> absolute (scratch dir on sys.path)
  * `a / b` on ints that py2 floor-divides    -> `a // b`  (only the cited sites)
  * float array sizes / slice bounds          -> int(...)  (old numpy truncated)
  * `np.complex`, `np.linalg.linalg`          -> `complex`, `np.linalg`
  * `range(n).remove(..)`                     -> `list(range(n))`
  * `unicode`, `string.join`                  -> py3 names
Modules that the hot path never calls (demixTF, SeparateLeadStereoTF, minqt,
nsgt, spatial, sourcefilter) are replaced by empty stubs so `audioModel`
imports.
"""
import os
import re
import sys
import types

REF = os.environ.get("PYFASST_REFERENCE", "/root/reference")
SCRATCH = os.environ.get("PYFASST_REF_PY3", "/tmp/pyfasst_ref_py3")

_PRINT_RE = re.compile(r"^(\s*)print\b")


def _strip_prints(src):
    """Replace every py2 print statement (with its continuation lines) by `pass`."""
    out = []
    lines = src.split("\n")
    i = 0
    while i < len(lines):
        line = lines[i]
        m = _PRINT_RE.match(line)
        if not m or line.lstrip().startswith("#"):
            out.append(line)
            i += 1
            continue
        indent = m.group(1)
        # swallow continuation lines: trailing backslash or unbalanced brackets
        stmt = line
        while True:
            code = re.sub(r"(\"(\\.|[^\"\\])*\"|'(\\.|[^'\\])*')", "", stmt)
            code = code.split("#")[0]
            depth = (code.count("(") - code.count(")") + code.count("[")
                     - code.count("]") + code.count("{") - code.count("}"))
            if stmt.rstrip().endswith("\\") or depth > 0:
                i += 1
                stmt = stmt + "\n" + lines[i]
                continue
            break
        # `if cond: print x` one-liners keep their prefix
        out.append(indent + "pass")
        i += 1
    return "\n".join(out)


def _strip_inline_prints(src):
    # `if self.verbose>1: print "..."` on one line
    return re.sub(r":[ \t]*print\b[^\n]*", ": pass", src)


# (file, [(old, new), ...]) -- literal replacements, every `old` must be present.
PATCHES = {
    "audioModel.py": [
        ("import audioObject as ao\nimport demixTF as demix\n",
         "import audioObject as ao\nimport demixTF as demix\nunicode = str\n"),
        # py2 integer division (audioModel.py:233,240,293)
        ("self.sig_repr_params['wlen']/2,", "self.sig_repr_params['wlen']//2,"),
        ("np.zeros(self.sig_repr_params['fsize']/2+1)",
         "np.zeros(self.sig_repr_params['fsize']//2+1)"),
        ("self.Cx = np.zeros([nc * (nc + 1) / 2,",
         "self.Cx = np.zeros([nc * (nc + 1) // 2,"),
        # range().remove (audioModel.py:1511)
        ("other_fact_ind_arr = range(nbfactors)",
         "other_fact_ind_arr = list(range(nbfactors))"),
        ("np.linalg.linalg.LinAlgError", "np.linalg.LinAlgError"),
    ],
    "audioObject.py": [
        ("from tools.utils import *", "from tools.utils import *"),
    ],
    "tftransforms/stft.py": [
        ("from ..tools.utils import *", "from tools.utils import *"),
        # float sizes / indices that old numpy truncated (stft.py:40-63)
        ("numberFrames = np.ceil(lengthData / np.double(hopsize)) + 2",
         "numberFrames = int(np.ceil(lengthData / np.double(hopsize)) + 2)"),
        ("data = np.concatenate((np.zeros(lengthWindow/2.0), data))",
         "data = np.concatenate((np.zeros(int(lengthWindow/2.0)), data))"),
        ("data = np.concatenate((data, np.zeros(newLengthData - data.size)))",
         "data = np.concatenate((data, np.zeros(int(newLengthData - data.size))))"),
        ("    numberFrequencies = nfft / 2 + 1\n    \n    STFT = np.zeros(",
         "    numberFrequencies = nfft // 2 + 1\n    \n    STFT = np.zeros("),
        # istft (stft.py:123-124)
        ("data = data[(lengthWindow/2.0):]", "data = data[int(lengthWindow/2.0):]"),
        ("normalisationSeq = normalisationSeq[(lengthWindow/2.0):]\n"
         "    normalisationSeq[normalisationSeq==0] = 1.\n"
         "    # ...added in the stft computation\n"
         "    \n"
         "    # normalising the liutkus way:\n"
         "    data = data / normalisationSeq",
         "normalisationSeq = normalisationSeq[int(lengthWindow/2.0):]\n"
         "    normalisationSeq[normalisationSeq==0] = 1.\n"
         "    # ...added in the stft computation\n"
         "    \n"
         "    # normalising the liutkus way:\n"
         "    data = data / normalisationSeq"),
        ("self.freqbins = self.ftlen / 2 + 1", "self.freqbins = self.ftlen // 2 + 1"),
        # filter_stft (stft.py:133-227): float sizes / slice bounds that old numpy truncated
        ("    numberFrames = np.ceil(lengthData / np.double(hopsize))\n    # to ensure that the data array s big enough,\n"
         "    # assuming the first frame is centered on first sample:\n"
         "    newLengthData = (numberFrames-1) * hopsize + lengthWindow\n    \n"
         "    # !!! adding zeros to the beginning of data, such that the first window is\n"
         "    # centered on the first sample of data\n"
         "    data = np.concatenate((np.zeros([lengthWindow/2.0, nc]), data))",
         "    numberFrames = int(np.ceil(lengthData / np.double(hopsize)))\n"
         "    newLengthData = int((numberFrames-1) * hopsize + lengthWindow)\n"
         "    data = np.concatenate((np.zeros([int(lengthWindow/2.0), nc]), data))"),
        ("    numberFrequencies = nfft / 2 + 1\n    if numberFrequencies != W.shape[2]:",
         "    numberFrequencies = int(nfft) // 2 + 1\n    if numberFrequencies != W.shape[2]:"),
        ("    for n in np.arange(numberFrames):\n        beginFrame = n * hopsize\n"
         "        endFrame = beginFrame + lengthWindow\n        \n        # Compute Fourier transforms",
         "    for n in np.arange(numberFrames):\n        beginFrame = int(n * hopsize)\n"
         "        endFrame = beginFrame + lengthWindow\n        \n        # Compute Fourier transforms"),
        ("    ndata = ndata[(lengthWindow/2.0):]\n    normalisationSeq = normalisationSeq[(lengthWindow/2.0):]",
         "    ndata = ndata[int(lengthWindow/2.0):]\n    normalisationSeq = normalisationSeq[int(lengthWindow/2.0):]"),
    ],
    "tools/utils.py": [
        # scipy moved the window functions to scipy.signal.windows (same function)
        ("spsig.blackmanharris(M)", "spsig.windows.blackmanharris(M)"),
    ],
    "tools/signalTools.py": [],
    "tools/nmf.py": [],
    "SeparateLeadStereo/SIMM/SIMM.py": [
        ("from string import join\n", ""),
    ],
    # generateHannBasis (the smooth-filter dictionary WGAMMA, SeparateLeadStereoTF.py:506-512)
    "sourcefilter/filter.py": [
        ("        lengthSineWindow = 2.0 * np.floor(lengthSineWindow / 2.0) ",
         "        lengthSineWindow = int(2.0 * np.floor(lengthSineWindow / 2.0))"),
        ("        sizeBigWindow = 2.0 * numberFrequencyBins",
         "        sizeBigWindow = 2 * int(numberFrequencyBins)"),
        ("    bigWindow[(sizeBigWindow - lengthSineWindow / 2.0):\\\n"
         "              (sizeBigWindow + lengthSineWindow / 2.0)] \\\n",
         "    bigWindow[(sizeBigWindow - lengthSineWindow // 2):\\\n"
         "              (sizeBigWindow + lengthSineWindow // 2)] \\\n"),
    ],
    # only stft / istft / sinebell of this module are executed (the SIMM front and back end,
    # SeparateLeadStereoTF.py:761-917, :1762-1871); the other transforms are stubbed
    "SeparateLeadStereo/separateLeadFunctions.py": [
        ("from ..tftransforms import minqt\nfrom ..tftransforms import nsgt\n"
         "from .. import audioObject as ao # for all these fancy transforms\n",
         "minqt = nsgt = ao = None\n"),
        ("from ..tools.utils import *\nfrom ..tools.distances import ISDistortion\n",
         "from tools.utils import *\n"),
        # float sizes / slice bounds that old numpy truncated (separateLeadFunctions.py:127-151)
        ("    data = np.concatenate((np.zeros(lengthWindow / 2.0),\n"
         "                           data,\n"
         "                           np.zeros(lengthWindow / 2.0)))",
         "    data = np.concatenate((np.zeros(int(lengthWindow / 2.0)),\n"
         "                           data,\n"
         "                           np.zeros(int(lengthWindow / 2.0))))"),
        ("    numberFrames = np.ceil((lengthData - lengthWindow) / hopsize \\\n"
         "                           + 1) + 1  ",
         "    numberFrames = int(np.ceil((lengthData - lengthWindow) / hopsize \\\n"
         "                           + 1) + 1)"),
        ("    data = np.concatenate((data, np.zeros([newLengthData - lengthData])))",
         "    data = np.concatenate((data, np.zeros([int(newLengthData - lengthData)])))"),
        ("    numberFrequencies = nfft / 2.0 + 1\n    \n    if stop is None:",
         "    numberFrequencies = int(nfft / 2.0 + 1)\n    \n    if stop is None:"),
        # the F0 dictionary generators (generate_WF0_chirped :237-345, generate_WF0_TR_chirped
        # :696-886, generate_ODGD_spec :888-949): float count that old numpy truncated when it
        # was used as a size / index (all the generators share the line)
        ("numberOfF0 = np.ceil(12.0 * stepNotes * np.log2(maxF0 / minF0)) + 1",
         "numberOfF0 = int(np.ceil(12.0 * stepNotes * np.log2(maxF0 / minF0)) + 1)"),
        # `array == 'sinebell'` was a scalar False in old numpy (the window may be an array)
        ("    if analysisWindowType=='sinebell':\n        analysisWindow = sinebell(lengthOdgd)\n"
         "    elif analysisWindowType=='hanning' or \\\n             analysisWindowType=='hanning':",
         "    if isinstance(analysisWindowType, str) and analysisWindowType=='sinebell':\n"
         "        analysisWindow = sinebell(lengthOdgd)\n"
         "    elif isinstance(analysisWindowType, str) and analysisWindowType=='hanning':"),
        ("    elif analysisWindowType=='rectangular':\n        analysisWindow = np.ones(lengthOdgd)\n"
         "    elif len(analysisWindowType)==lengthOdgd:",
         "    elif isinstance(analysisWindowType, str) and analysisWindowType=='rectangular':\n"
         "        analysisWindow = np.ones(lengthOdgd)\n"
         "    elif len(analysisWindowType)==lengthOdgd:"),
        # old np.fft.rfft cast a complex input to float (ComplexWarning: the imaginary part is
        # discarded); today's raises.  The STFT object receives the complex waveform (:846, :877)
        # np.load of the .npz cache (it holds the pickled transform object): pickles were allowed
        # by default in the NumPy of the time
        ("struc = np.load(filename)", "struc = np.load(filename, allow_pickle=True)"),
        ("transform.computeTransform(data=odgd)\n",
         "transform.computeTransform(data=np.real(odgd))\n"),
    ],
}

_GLOBAL_SUBS = [
    (re.compile(r"\bnp\.complex\b(?!\d|_)"), "complex"),
    (re.compile(r"\bnp\.NaN\b"), "np.nan"),
    (re.compile(r"\bnp\.int\b(?!\d|_|e)"), "int"),
    (re.compile(r"\bnp\.float\b(?!\d|_|i)"), "float"),
]


def _patch(rel, src):
    src = _strip_inline_prints(src)
    src = _strip_prints(src)
    for old, new in PATCHES[rel]:
        if old not in src:
            raise RuntimeError("patch site not found in %s: %r" % (rel, old))
        src = src.replace(old, new)
    for rx, new in _GLOBAL_SUBS:
        src = rx.sub(new, src)
    return src


def build(scratch=SCRATCH):
    """Write the patched copies; return the scratch dir (to put on sys.path)."""
    for rel in PATCHES:
        with open(os.path.join(REF, "pyfasst", rel)) as fh:
            src = fh.read()
        dst = os.path.join(scratch, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as fh:
            fh.write(_patch(rel, src))
    for pkg in ("tftransforms", "tools", "SeparateLeadStereo",
                "SeparateLeadStereo/SIMM", "sourcefilter"):
        open(os.path.join(scratch, pkg, "__init__.py"), "w").close()
    # the TF-transform registry: keep only the STFT entry (tft.py:74-80)
    with open(os.path.join(scratch, "tftransforms", "tft.py"), "w") as fh:
        fh.write("from tftransforms.stft import STFT\n"
                 "class TFTransform(object):\n    pass\n"
                 "MinQTransfo = CQTransfo = NSGMinQT = TFTransform\n")
    return scratch


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load():
    """Import the patched reference; returns (audioModel, stft module, SIMM module)."""
    scratch = build()
    if scratch not in sys.path:
        sys.path.insert(0, scratch)
    _stub("demixTF")
    sls = _stub("SeparateLeadStereo.SeparateLeadStereoTF")
    import SeparateLeadStereo
    SeparateLeadStereo.SeparateLeadStereoTF = sls
    _stub("spatial")
    _stub("spatial.steering_vectors",
          gen_steer_vec_far_src_uniform_linear_array=None)
    import audioModel
    import tftransforms.stft as ref_stft
    import SeparateLeadStereo.SIMM.SIMM as ref_simm
    import tools.signalTools as ref_st
    import tools.utils as ref_utils
    import tools.nmf as ref_nmf
    import SeparateLeadStereo.separateLeadFunctions as ref_slf
    import sourcefilter.filter as ref_filter
    return dict(audioModel=audioModel, stft=ref_stft, SIMM=ref_simm,
                signalTools=ref_st, utils=ref_utils, nmf=ref_nmf, slf=ref_slf, filter=ref_filter,
                writeSeparatedSignals=_method_source(
                    "SeparateLeadStereo/SeparateLeadStereoTF.py", "writeSeparatedSignals"),
                runViterbi=_method_source(
                    "SeparateLeadStereo/SeparateLeadStereoTF.py", "runViterbi"),
                initiateHF0WithIndexBestPath=_method_source(
                    "SeparateLeadStereo/SeparateLeadStereoTF.py", "initiateHF0WithIndexBestPath"))


class OldNumpy(object):
    """`np` as the reference's Python-2-era code expects it, for exec()ing method sources:
    removed aliases (np.Inf) and float array sizes (old NumPy truncated them with a warning).
    Everything else is today's NumPy."""

    def __init__(self):
        import numpy
        self._np = numpy
        self.Inf = numpy.inf

    def __getattr__(self, name):
        return getattr(self._np, name)

    def _size(self, shape):
        if isinstance(shape, (list, tuple)):
            return [int(v) for v in shape]
        return int(shape)

    def ones(self, shape, *a, **k):
        return self._np.ones(self._size(shape), *a, **k)

    def zeros(self, shape, *a, **k):
        return self._np.zeros(self._size(shape), *a, **k)


def _method_source(rel, name):
    """The (print-stripped, dedented) source of one method of a reference class, for modules
    that cannot be imported as a whole (SeparateLeadStereoTF needs the compiled Viterbi
    extension): the caller exec()s it against a stand-in `self`."""
    import textwrap
    with open(os.path.join(REF, "pyfasst", rel)) as fh:
        src = fh.read()
    m = re.search(r"^    def %s\(.*?(?=^    def )" % name, src, re.S | re.M)
    if m is None:
        raise RuntimeError("method %s not found in %s" % (name, rel))
    return textwrap.dedent(_strip_prints(_strip_inline_prints(m.group(0))))


if __name__ == "__main__":
    mods = load()
    print("loaded:", sorted(mods))
