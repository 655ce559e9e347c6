"""Import the UNMODIFIED reference modules from /root/reference -- TEST INFRASTRUCTURE ONLY.

Works in the build container only (the GPU box has no /root/reference); used by
oracle/make_golden.py and by the CPU tests that are skipped when the tree is absent.

Four shims make the reference importable on this image:
  * ``kornia``             -> oracle.kornia050 (restated 0.5.0 functions)
  * ``matplotlib.pyplot``  -> empty stub (only src/data/coco/dataset.py:5 imports it)
  * ``torchvision.models.resnet*(pretrained=True)`` -> seeded random init (no network);
    hit by src/heads/PerceptualHead.py:22.
  * ``torch.hub._download_url_to_file`` -> today's public name (src/utils/model_zoo.py:9-16; never called).
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('BIHOME_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'src', 'heads', 'PerceptualHead.py'))


def _install_shims():
    from . import kornia050
    sys.modules['kornia'] = kornia050
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:  # noqa: BLE001
            mpl = types.ModuleType('matplotlib')
            plt = types.ModuleType('matplotlib.pyplot')
            mpl.pyplot = plt
            sys.modules['matplotlib'] = mpl
            sys.modules['matplotlib.pyplot'] = plt
    import torch
    import torch.hub
    import torchvision.models as tvm
    # src/utils/model_zoo.py:9-16 imports a private downloader that newer torch renamed; nothing here downloads
    for name, value in (('_download_url_to_file', torch.hub.download_url_to_file),):
        if not hasattr(torch.hub, name):
            setattr(torch.hub, name, value)
    if not getattr(tvm, '_bihome_oracle_patched', False):
        for name in ('resnet18', 'resnet34', 'resnet50'):
            orig = getattr(tvm, name)

            def make(orig_fn):
                def build(pretrained=False, progress=True, **kw):
                    state = torch.random.get_rng_state()
                    torch.manual_seed(1234)
                    kw.pop('weights', None)
                    try:
                        return orig_fn(weights=None, **kw)
                    finally:
                        torch.random.set_rng_state(state)
                return build
            setattr(tvm, name, make(orig))
        tvm._bihome_oracle_patched = True


def load(module_name):
    """e.g. load('src.heads.PerceptualHead') -> the reference's module object."""
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    mod = importlib.import_module(module_name)
    assert os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), mod.__file__
    return mod
