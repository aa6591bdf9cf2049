"""`stylegan2ada.torch_utils.ops` served by sgb200 (B200-native kernels).  Importing this package installs the
op modules under the reference's dotted names and the post-import hook that rebinds `modulated_conv2d`."""
import os
import sys

_pkg_root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..', '..', '..'))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)

import sgb200  # noqa: E402

for _name, _mod in sgb200.install().items():
    globals()[_name.rsplit('.', 1)[1]] = _mod
