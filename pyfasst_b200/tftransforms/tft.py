"""Time-frequency transform registry (ref: pyfasst/tftransforms/tft.py:12-80).

Only the STFT front end is on the accelerated path (BASELINE.json north_star); the
constant-Q family of the reference (minqt / cqt / nsgmqt) is out of scope and is not
registered, so `FASST(transf='mqt')` raises NotImplementedError like an unknown name
does in the reference (audioModel.py:201-204).
"""
from .stft import STFT


class TFTransform(object):
    """Duck-type contract of a transform: computeTransform(data) fills `.transfo`,
    invertTransform() inverts it (ref: tft.py:12-71)."""
    transformname = 'dummy'
    transfo = None

    def __init__(self, **kwargs):
        pass

    def computeTransform(self, data):
        pass

    def invertTransform(self):
        pass


tftransforms = {'stft': STFT}
