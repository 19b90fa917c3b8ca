"""Host-side mirror of the reference's `modules/` package for the fusion hot path.

Same class names, constructor arguments, forward signatures and return values as
modules/{extractor,integrator,pipeline,model,adapnet,database}.py of the reference, so
`from modules.pipeline import Pipeline` can be re-pointed at this package
(see INTEGRATION.md).  The work itself is done by csrc/libojdf.so.
"""
from .extractor import Extractor  # noqa: F401
from .integrator import Integrator  # noqa: F401
