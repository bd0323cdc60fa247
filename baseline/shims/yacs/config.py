from linnaeus_b200.config import CfgNode  # noqa: F401  (yacs-compatible attribute dict: get / clone / defrost / freeze / merge_*)
