"""Run the reference's own scripts UNCHANGED on the B200 kernels.

The reference has no plugin layer: its scripts import ``fact_clip.models.blocks`` directly (scripts/train.py:192,199,202;
scripts/run_eval.py:124-128; scripts/eval.py:26-27; scripts/fact_input_emb_logit_viz.py:16).  ``install()`` puts an import
hook in front of the normal finders that resolves exactly those model modules to this package, while every other
``fact_clip.*`` module (configs, dataset, evaluate, train_tools, text_embeddings, loss) keeps coming from the installed
reference:

    import fact_clip_b200.dropin            # installs the hook (idempotent)
    from fact_clip.models.blocks import FACT_CLIP      # -> fact_clip_b200.models.blocks.FACT_CLIP

or, without touching a script at all:

    python -m fact_clip_b200.dropin scripts/run_eval.py --cfg ... --ckpt ...

``FACTK_MODE=fp32|bf16`` selects the compute mode of models built that way (default bf16).
"""
import importlib
import importlib.abc
import importlib.util
import runpy
import sys

ALIASES = {
    'fact_clip.models.blocks': 'fact_clip_b200.models.blocks',
    'fact_clip.models.blocks_SepVerbNoun': 'fact_clip_b200.models.blocks_SepVerbNoun',
}


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in ALIASES:
            return importlib.util.spec_from_loader(fullname, self, origin=ALIASES[fullname])
        return None

    def create_module(self, spec):
        return None         # a fresh module object; exec_module fills it

    def exec_module(self, module):
        real = importlib.import_module(ALIASES[module.__name__])
        for k, v in vars(real).items():
            if not (k.startswith('__') and k.endswith('__')):
                setattr(module, k, v)
        module.__factk_real__ = real


_finder = None


def install():
    global _finder
    if _finder is None:
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)
        for name in ALIASES:              # a copy imported before the hook existed would shadow it
            sys.modules.pop(name, None)
    return _finder


def uninstall():
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        for name in ALIASES:
            sys.modules.pop(name, None)
        _finder = None


install()


if __name__ == '__main__':
    if len(sys.argv) < 2:
        sys.exit('usage: python -m fact_clip_b200.dropin <reference script.py> [script arguments ...]')
    sys.argv = sys.argv[1:]
    runpy.run_path(sys.argv[0], run_name='__main__')
