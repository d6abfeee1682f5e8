"""`pyfasst` -- the reference's package name, resolving to the B200 implementation.

A user of s-ben/pyfasst keeps their imports for the accelerated path
(doc/source/description.rst:50-100 of the reference):

    import pyfasst.audioModel as am
    model = am.MultiChanNMFInst_FASST(audio="mix.wav", nbComps=3, iter_num=50)
    from pyfasst.SeparateLeadStereo.SIMM.SIMM import Stereo_SIMM
    from pyfasst.tftransforms.stft import stft, istft, filter_stft

Every name below is the SAME module object as its `pyfasst_b200` namesake (no second copy of
the code or of its state).  Modules of the reference that are out of scope (demixTF, minqt,
nsgt, spatial, ... see DESIGN.md section 8) are not aliased: importing them raises ImportError.
"""
import importlib
import sys

_ALIASES = (
    "audioModel",
    "audioObject",
    "tftransforms",
    "tftransforms.stft",
    "tftransforms.tft",
    "tools",
    "tools.utils",
    "tools.nmf",
    "SeparateLeadStereo",
    "SeparateLeadStereo.separateLeadFunctions",
    "SeparateLeadStereo.SeparateLeadStereoTF",
    "SeparateLeadStereo.SIMM",
    "SeparateLeadStereo.SIMM.SIMM",
    "SeparateLeadStereo.tracking",
    "SeparateLeadStereo.tracking._tracking",
)

for _name in _ALIASES:
    _mod = importlib.import_module("pyfasst_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    _parent, _, _leaf = (__name__ + "." + _name).rpartition(".")
    setattr(sys.modules[_parent], _leaf, _mod)

del importlib, sys, _name, _mod, _parent, _leaf
