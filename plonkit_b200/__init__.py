"""plonkit_b200 — B200-native prove path for fluidex/plonkit (PLONK over BN254).

Host-side mirror of the reference's API for the prove path (src/plonk.rs) over a C-ABI CUDA library
(include/plonkit_b200.h, built from plonkit_b200/csrc/).  Importing the package does not load the library;
anything that computes does, and fails loudly when it is missing (there is no CPU fallback).
"""
__version__ = "0.1.0"
