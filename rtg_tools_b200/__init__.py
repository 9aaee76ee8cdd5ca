"""B200-native vcfeval matching engine (drop-in for the hot path of RealTimeGenomics/rtg-tools `vcfeval`).

Only what the path needs lives here: csrc/ (CUDA kernels + the C-ABI of include/vce.h), capi.py (ctypes view
of that ABI), vcfeval.py (host-side mirror of the reference's SequenceEvaluator / PathFinder / RocContainer
interface) and synth.py (seeded synthetic inputs for tests and bench.py).
"""
__version__ = "0.1.0"
