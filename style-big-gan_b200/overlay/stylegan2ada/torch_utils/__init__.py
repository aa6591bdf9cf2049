"""Path overlay: put `style-big-gan_b200/overlay` AHEAD of the reference checkout on PYTHONPATH.

`stylegan2ada` itself stays a namespace package (the reference has no stylegan2ada/__init__.py), so its other
sub-packages (dnnlib, training, metrics) still come from the reference.  This package shadows
`stylegan2ada.torch_utils` and re-exposes the reference's own modules (misc, persistence, training_stats,
custom_ops) by appending the reference's directory to __path__; only `ops` is replaced."""
import os
import sys

for _p in sys.path:
    _cand = os.path.join(_p, 'stylegan2ada', 'torch_utils')
    if os.path.isdir(_cand) and os.path.abspath(_cand) != os.path.dirname(os.path.abspath(__file__)):
        __path__.append(_cand)
        break
