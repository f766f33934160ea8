"""Import alias: `import topicgcn_b200` -> graph-convolutional-networks-for-text-classification_b200/.

The product package directory carries the reference repo's name (with hyphens, not a Python identifier); this shim
points the importable name at it so `topicgcn_b200.layer`, `topicgcn_b200.ops`, ... resolve to those files.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "graph-convolutional-networks-for-text-classification_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
