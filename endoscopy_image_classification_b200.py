"""Import shim: exposes the directory ``endoscopy-image-classification_b200/`` (a
hyphen is not a valid identifier) as the package ``endoscopy_image_classification_b200``."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "endoscopy-image-classification_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f
