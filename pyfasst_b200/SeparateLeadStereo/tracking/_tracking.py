"""viterbiTracking: drop-in for the reference's one native module,
pyfasst/SeparateLeadStereo/tracking/_tracking.pyx (:11-93, Cython).  Same signature, NumPy in /
NumPy out; the recursion runs on the GPU (csrc/viterbi.cu) in float64 with the reference's tie
breaking, so the decoded path is identical.  No CPU fallback."""
import numpy as np


def viterbiTracking(numberOfStates, numberOfFrames, logDensity, logPriorDensities,
                    logTransitionMatrix, verbose=False, kernels=None):
    """bestStatePath = viterbiTracking(S, N, logDensity[S', N'], logPriorDensities[S'],
    logTransitionMatrix[S', S']): the most likely state sequence of the HMM.  Like the reference,
    only the first `numberOfStates` states and `numberOfFrames` frames of the arrays are used
    (SeparateLeadStereoTF.py:1217-1219 passes arrays with one more state than it decodes)."""
    import torch
    if kernels is None:
        from ...tftransforms.stft import default_kernels
        kernels = default_kernels()
    S, N = int(numberOfStates), int(numberOfFrames)
    dens = np.ascontiguousarray(np.asarray(logDensity, dtype=np.float64)[:S, :N])
    prior = np.ascontiguousarray(np.asarray(logPriorDensities, dtype=np.float64)[:S])
    trans = np.ascontiguousarray(np.asarray(logTransitionMatrix, dtype=np.float64)[:S, :S])
    dev = kernels.device
    path = kernels.viterbi(torch.from_numpy(dens).to(dev), torch.from_numpy(prior).to(dev),
                           torch.from_numpy(trans).to(dev))
    return path.cpu().numpy()
