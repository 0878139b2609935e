"""cocodr_b200: B200-native (sm_100a) implementation of the COCO-DR contrastive hot path.

Host side mirrors the reference's model surface (ANCE/model/models.py, COCO/modeling.py,
evaluate/ scan); every compute op goes through the C ABI in include/cocodr_b200.h.  There is no CPU
or eager-PyTorch fallback: importing the ops without the built CUDA library raises.
"""
__version__ = "0.1.0"
