"""Drop-in alias: `import SPART` resolves to the B200 implementation with the reference's
module-level names (reference src/SPART/__init__.py:1-5)."""
from spart_b200 import *          # noqa: F401,F403
from spart_b200 import SPART, run_batch, run_batch_params  # noqa: F401
