"""B200-native Asynchronous-Score-Distillation step behind threestudio's plugin API.

`import scaledreamer_b200` registers the plugins under the reference's names; the arithmetic
lives in libsdb200.so (csrc/, C ABI in include/*.h). There is no CPU or PyTorch fallback for the CUDA path.
"""
__version__ = "0.1.0"

from .core import C, find, load_config, parse_structured, register  # noqa: F401
from . import data, fields, guidance, prompts, systems, amortized  # noqa: F401  (registration side effects)
