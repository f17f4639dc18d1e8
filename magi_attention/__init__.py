"""Drop-in alias package: ``import magi_attention`` resolves to the B200-native engine.

The reference's examples and integrations import ``magi_attention.api`` / ``magi_attention.functional`` /
``magi_attention.dist_attn_runtime_mgr`` (examples/torch_native/main.py:28-36, modeling_llama.py:24); these thin
modules re-export the same names from ``magiattention_b200`` so that user code runs unchanged.  Nothing is
implemented here.
"""
from magiattention_b200 import __version__  # noqa: F401
from magiattention_b200 import config  # noqa: F401

from . import api, comm, common, functional  # noqa: F401,E402
