from .inference import PostProcessor, box_postprocess, make_roi_box_post_processor  # noqa: F401
