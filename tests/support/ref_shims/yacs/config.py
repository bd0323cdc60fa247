"""TEST SHIM: `yacs.config.CfgNode` backed by linnaeus_b200.config.CfgNode."""
from linnaeus_b200.config import CfgNode  # noqa: F401
