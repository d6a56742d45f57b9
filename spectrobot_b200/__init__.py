"""spectrobot_b200: B200-native line-by-line forward model behind the SpectRobot Python API."""
import importlib
import sys

REFERENCE_MODULES = ('spect_base_module', 'spect_classes', 'spect_main_module', 'lineshape',
                     'fparts_mod', 'curgods')


def install_reference_names():
    """Registers this package's modules under the reference's top-level module names, so that a
    driver written against the reference (`import spect_classes as spcl`, `import
    spect_main_module as smm`, `import spect_base_module as sbm`, `import lineshape`, ...) runs on
    this implementation unchanged."""
    for name in REFERENCE_MODULES:
        sys.modules[name] = importlib.import_module(__name__ + '.' + name)
