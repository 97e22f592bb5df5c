"""inconsistencymasks_b200 -- B200-native (sm_100a) pseudo-label generation hot path of
MichaelVorndran/InconsistencyMasks: U-Net ensemble forward -> threshold / argmax ->
Inconsistency Mask -> blanking, behind the reference's own Python helper API.

    from inconsistencymasks_b200 import functions, unet

The compute lives in ``libimk.so`` (hand-written CUDA behind the C ABI of include/imk.h);
importing ``functions`` / ``unet`` fails loudly when the library has not been built.
"""
__version__ = "0.1.0"
