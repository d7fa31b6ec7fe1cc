"""Import the UNMODIFIED reference leaf modules from /root/reference (container only).

Recipe: the reference's package ``__init__`` files pull in pytorch_lightning /
omegaconf / hydra / h5py, none of which are installed.  We therefore (1) register stub ``omegaconf`` and
``h5py`` modules, (2) pre-register empty namespace packages for the parent packages so their heavy
``__init__.py`` never runs, (3) import only the leaf modules on the hot path.  Nothing is copied.
Used by ``oracle/make_golden.py`` and by tests that are skipped when /root/reference is absent.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("MRIDC_REFERENCE_ROOT", "/root/reference")

_PKGS = [
    "mridc",
    "mridc.collections",
    "mridc.collections.common",
    "mridc.collections.common.parts",
    "mridc.collections.common.data",
    "mridc.collections.reconstruction",
    "mridc.collections.reconstruction.data",
    "mridc.collections.reconstruction.models",
    "mridc.collections.reconstruction.models.rim",
    "mridc.collections.reconstruction.models.varnet",
    "mridc.collections.reconstruction.models.unet_base",
    "mridc.collections.quantitative",
    "mridc.collections.quantitative.models",
    "mridc.collections.quantitative.models.qrim",
    "mridc.collections.reconstruction.models.sigmanet",
    "mridc.collections.reconstruction.models.cascadenet",
    "mridc.collections.reconstruction.models.recurrentvarnet",
    "mridc.collections.quantitative.models.qvarnet",
    "mridc.collections.reconstruction.parts",
    "mridc.collections.segmentation",
    "mridc.collections.segmentation.models",
    "mridc.collections.segmentation.models.jrscirim_base",
    "mridc.collections.segmentation.models.attention_unet_base",
    "mridc.collections.segmentation.models.lambda_unet_base",
    "mridc.collections.segmentation.models.vnet_base",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "mridc", "collections"))


def _setup():
    if "mridc" in sys.modules and getattr(sys.modules["mridc"], "__oracle_stub__", False):
        return
    sys.dont_write_bytecode = True  # never write __pycache__ into the read-only reference
    here = os.path.dirname(os.path.abspath(__file__))
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(here, "_ref", "numba"))
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")

        class ListConfig(list):
            pass

        class DictConfig(dict):
            pass

        oc.ListConfig = ListConfig
        oc.DictConfig = DictConfig
        sys.modules["omegaconf"] = oc
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    for name in _PKGS:
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
        mod.__oracle_stub__ = True
        sys.modules[name] = mod


def ref(name: str):
    """Import e.g. ref('common.parts.fft') -> mridc.collections.common.parts.fft (reference code)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _setup()
    return importlib.import_module("mridc.collections." + name)


class Ref:
    """Lazy bundle of the hot-path leaf modules."""

    def __init__(self):
        self.fft = ref("common.parts.fft")
        self.utils = ref("common.parts.utils")
        self.subsample = ref("reconstruction.data.subsample")
        self.subsample_nn = ref("common.data.subsample")
        self.rim_utils = ref("reconstruction.models.rim.rim_utils")
        self.rim_block = ref("reconstruction.models.rim.rim_block")
        self.rnn_cells = ref("reconstruction.models.rim.rnn_cells")
        self.conv_layers = ref("reconstruction.models.rim.conv_layers")
        self.vn_block = ref("reconstruction.models.varnet.vn_block")
        self.unet_block = ref("reconstruction.models.unet_base.unet_block")
        self.qrim_utils = ref("quantitative.models.qrim.utils")
        self.qrim_block = ref("quantitative.models.qrim.qrim_block")

    # the other consumers of the DC operator, imported on first use
    @property
    def dc_layers(self):
        return ref("reconstruction.models.sigmanet.dc_layers")

    @property
    def ccnn_block(self):
        return ref("reconstruction.models.cascadenet.ccnn_block")

    @property
    def recurrentvarnet(self):
        return ref("reconstruction.models.recurrentvarnet.recurrentvarnet")

    @property
    def jrscirim_block(self):
        return ref("segmentation.models.jrscirim_base.jrscirim_block")

    @property
    def transforms(self):
        return ref("reconstruction.parts.transforms")

    @property
    def qvn_block(self):
        return ref("quantitative.models.qvarnet.qvn_block")


def ref_class_from_source(rel_path: str, class_name: str, namespace: dict):
    """Execute ONE class definition of a reference file that cannot be imported as a module (its imports need
    pytorch_lightning / omegaconf): the class node is cut out of the parsed source and compiled with the file's own
    name, in a namespace holding the (importable) reference leaf modules it uses.  Nothing is copied into the repo."""
    import ast

    path = os.path.join(REF_ROOT, rel_path)
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), namespace)
            return namespace[class_name]
    raise LookupError("%s not found in %s" % (class_name, path))
