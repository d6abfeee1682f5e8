"""WAV input/output with the reference's scaling conventions (host side).

Mirrors pyfasst/audioObject.py:76-195 (the scipy.io.wavfile branch): samples are
divided by 1.1 * max|x| on read (`_maxdata`, :124-127), `data` is a lazy property
whose deleter forces a re-read, and writing multiplies back by `_maxdata`.
"""
import warnings

import numpy as np
import scipy.io.wavfile as wav

from .tools.utils import *  # noqa: F401,F403  (the reference re-exports these)


def wavread(filename, first=0, last=None):
    """ref: audioObject.py:76-81"""
    fs, data = wav.read(filename)
    data = data[first:last]
    return fs, data, data.dtype


def wavwrite(filename, rate, data, formattype='wav', formatenc='int16', formatend='file'):
    """ref: audioObject.py:83-98 -- the encoding is re-derived from the peak value
    when it is not one of the integer names."""
    if formatenc not in ('int16', 'int32', 'int8'):
        peak = np.abs(data).max() if np.size(data) else 0
        if peak > 2 ** 15:
            formatenc = 'int32'
        elif peak > 2 ** 7:
            formatenc = 'int16'
        else:
            formatenc = 'int8'
    wav.write(filename, rate, np.array(data, dtype=formatenc))
    return 0


class AudioObject(object):
    """ref: audioObject.py:100-195"""

    def __init__(self, filename, mode='rw'):
        self.filename = filename
        self.mode = mode

    def _read(self):
        if 'r' not in self.mode:
            raise ValueError("Not in read mode.")
        self._samplerate, self._data, self._encoding = wavread(self.filename)
        if len(self._data.shape) == 2:
            self._nframes, self._channels = self._data.shape
        else:
            self._nframes = self._data.size
            self._channels = 1
        self._maxdata = np.maximum(1.1 * np.abs(self._data).max(), 1e-10)
        self._data = self._data / self._maxdata

    def _write(self):
        if 'w' not in self.mode:
            raise ValueError("Not in write mode.")
        if not hasattr(self, '_samplerate') and not hasattr(self, '_data'):
            raise AttributeError("Should set sample rate and have data in write mode.")
        wavwrite(filename=self.filename, rate=self._samplerate,
                 data=self._maxdata * self._data, formatenc=self._encoding)

    def _set_data(self, data):
        s = data.shape
        if s[0] < s[1] and s[1] > 2:
            self._data = np.array(data.T, order='C')
        else:
            self._data = np.array(data, order='C')
        self._maxdata = 1.1 * np.abs(self._data).max()
        self._encoding = self._data.dtype.name
        self._data = self._data / self._maxdata

    def _get_data(self):
        if not hasattr(self, '_data'):
            self._read()
        return self._data

    def _del_data(self):
        if hasattr(self, '_data'):
            del self._data

    data = property(_get_data, _set_data, _del_data)

    def _get_samplerate(self):
        if not hasattr(self, '_samplerate') and 'r' in self.mode:
            self._read()
        return self._samplerate

    def _set_samplerate(self, samplerate):
        if 'r' in self.mode:
            warnings.warn("Changing the sampling rate in read mode")
        self._samplerate = int(samplerate)

    samplerate = property(_get_samplerate, _set_samplerate)
    fs = samplerate

    @property
    def channels(self):
        if not hasattr(self, '_channels'):
            self._read()
        return self._channels

    @property
    def nframes(self):
        if not hasattr(self, '_nframes'):
            self._read()
        return self._nframes
