"""madigan_b200: B200-native batched trading environment behind Madigan's env API.

Only the Env.step hot path of ben-watt-es/madigan is implemented (price generators,
ledger, rewards, window observations) as hand-written sm_100a CUDA kernels reached
through the C-ABI of ``include/madigan_b200.h``.  Importing the package does not load
the CUDA library; constructing an ``Env`` does, and fails loudly if it is missing.
"""
__version__ = "0.1.0"
