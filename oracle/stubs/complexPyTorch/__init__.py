"""Restatement (TEST INFRASTRUCTURE ONLY) of the third-party package complexPyTorch==0.3.

The reference pins `complexPyTorch==0.3` (`/root/reference/requirements.txt:38`) and takes
ComplexConv2d / ComplexConvTranspose2d / ComplexBatchNorm2d / ComplexLinear / ComplexReLU /
complex_upsample from it (`c_network.py:5-7`, `config.py:5`).  The package is neither vendored
in the reference nor installed in this image, so its published semantics are restated here
 so that the reference's own files can run unmodified as the oracle.
PARITY UNPINNED: the reference ships no tests or golden vectors for this path; this
restatement is cross-checked only against the reference's call sites and state_dict shapes.
"""
