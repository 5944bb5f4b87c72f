"""TEST INFRASTRUCTURE ONLY -- restated subset of torch-geometric (pinned >=2.3,<2.5 by the
reference's pyproject.toml:49).  PyG is not installed in this image and cannot be fetched, so the
few entry points the reference's graph path calls are restated here from the published PyG API
(SURVEY.md section 8c lists the semantics).  Nothing under oracle/ is imported by the product
package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm use it.
"""
__version__ = "2.4.0+oracle.shim"
from . import data, nn, typing, utils  # noqa: F401,E402
