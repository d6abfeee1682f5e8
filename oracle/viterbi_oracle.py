"""CPU oracle for the Viterbi melody tracker.  TEST INFRASTRUCTURE ONLY.

NumPy restatement of pyfasst/SeparateLeadStereo/tracking/_tracking.pyx: viterbiTracking
(:11-93).  Parity status: PINNED -- tests/test_viterbi_cpu.py checks it against
tests/golden/viterbi.npz, produced by the reference's own Cython module compiled here
(oracle/build_ref.py -> oracle/_ref/), and against that module directly when it is present.
"""
import numpy as np


def viterbi_tracking(numberOfStates, numberOfFrames, logDensity, logPriorDensities,
                     logTransitionMatrix):
    S, N = numberOfStates, numberOfFrames
    dens = np.asarray(logDensity, dtype=np.float64)
    trans = np.asarray(logTransitionMatrix, dtype=np.float64)[:S, :S]
    cum = np.asarray(logPriorDensities, dtype=np.float64)[:S] + dens[:S, 0]   # :58-61
    ante = np.zeros([N, S], dtype=np.int64)
    for n in range(1, N):                                                      # :63-85
        cand = cum[:, None] + trans   # cand[s_, s] = cum[s_, n-1] + logT[s_, s]
        # the reference starts from s_ = 0 and replaces on a strict `>`: the first maximum
        ante[n] = np.argmax(cand, axis=0)
        cum = cand[ante[n], np.arange(S)] + dens[:S, n]
    path = np.zeros(N, dtype=np.int64)                                         # :88-92
    path[N - 1] = np.argmax(cum)
    for n in range(N - 2, -1, -1):
        path[n] = ante[n + 1, path[n + 1]]
    return path
