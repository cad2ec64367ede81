"""Import path of the reference's criterion (fact_clip/models/loss.py): ``from fact_clip_b200.models.loss import
MatchCriterion``.  The arithmetic lives in fact_clip_b200/loss.py (host orchestration) and csrc/loss.cu (kernels)."""
from ..loss import MatchCriterion  # noqa: F401
