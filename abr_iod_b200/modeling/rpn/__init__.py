from .inference import RPNPostProcessor, make_rpn_postprocessor, rpn_proposals  # noqa: F401
