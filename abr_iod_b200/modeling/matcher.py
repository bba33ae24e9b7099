"""``Matcher`` parameters (modeling/matcher.py:5-40 of the reference).  The matching itself runs inside
``abr_match_proposals``; any object with ``high_threshold``, ``low_threshold`` and ``allow_low_quality_matches`` --
including the reference's own Matcher -- can be passed where a proposal matcher is expected."""


class Matcher(object):
    BELOW_LOW_THRESHOLD = -1
    BETWEEN_THRESHOLDS = -2

    def __init__(self, high_threshold, low_threshold, allow_low_quality_matches=False):
        assert low_threshold <= high_threshold
        self.high_threshold = high_threshold
        self.low_threshold = low_threshold
        self.allow_low_quality_matches = allow_low_quality_matches
