"""Drop-in for /root/reference/models/deep_sets.py: put `point-cloud-classifier_b200/` ahead of
the reference checkout on sys.path and `from models.deep_sets import DeepSets` (train.py:9)
resolves here, while models.wrapper / train / sweep / utils still resolve to the reference
(models/ is a PEP 420 namespace package on both sides — no __init__.py)."""
from pcc_b200.deep_sets import DeepSets, ResidualBlock  # noqa: F401
