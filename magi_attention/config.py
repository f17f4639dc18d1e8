from magiattention_b200.config import *  # noqa: F401,F403
from magiattention_b200.config import __all__  # noqa: F401
