"""abr_iod_b200 -- B200-native (sm_100a) RoI hot path of ABR_IOD behind the reference's own Python API.

    from abr_iod_b200.layers import ROIAlign, ROIPool, nms, roi_align, roi_pool
    from abr_iod_b200.modeling.poolers import Pooler, LevelMapper, make_pooler
    from abr_iod_b200.structures.boxlist_ops import boxlist_nms, boxlist_nms_batched
    from abr_iod_b200.distillation.distillation import calculate_attentive_roi_feature_distillation
    from abr_iod_b200.data.abr_paste import BoxRehearsalPaster

``abr_iod_b200.compat.install()`` registers the same objects under ``maskrcnn_benchmark.*`` so that the
reference's training scripts pick them up unchanged.
"""
__version__ = "1.0.0"
