"""ec504_imageencoder_b200 -- B200-native (sm_100a) MPEG-1 I-frame per-block encode path behind
the reference's C API.  See DESIGN.md.  The compute path is csrc/ (hand-written CUDA) reached
through the C ABI in include/m1cu.h; this package is the thin Python host layer over it."""
from .encoder import (DEFAULT_QUALITY, MODE_FULL, MODE_REF_COMPAT, SYNTH_NATURAL, SYNTH_NOISE, SYNTH_GREY, SYNTH_RG_EQUAL, SYNTH_SCATTERED,
                      EncodedBatch, M1Encoder, M1Error, qmatrix)

__all__ = ["M1Encoder", "M1Error", "EncodedBatch", "qmatrix", "MODE_FULL", "MODE_REF_COMPAT",
           "SYNTH_NATURAL", "SYNTH_NOISE", "SYNTH_GREY", "SYNTH_RG_EQUAL", "SYNTH_SCATTERED", "DEFAULT_QUALITY"]
