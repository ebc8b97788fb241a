"""Loader of the UNMODIFIED reference installed under baseline/_ref (git-ignored; it travels to the GPU box).

The install is the base contract's recipe, run by ``__graft_entry__.build()`` when /root/reference is present:

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <a /tmp copy of /root/reference>

The reference's setup.py installs its packages top-level (``model``, ``multimodal``, ``training`` ...) while its own
modules import each other as ``src.<package>`` (core.py:64, pipeline.py:21: it is meant to be run from its repo
root).  Nothing is edited: ``src`` is registered here as a namespace alias whose search path is the install
directory, so ``src.model.core`` resolves to ``baseline/_ref/model/core.py`` byte for byte as upstream ships it.

Only tests/, bench.py (the reference arm and the cpu_baseline leg) and tools/ import this; the product package
``apertis_llm_b200`` never does.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "core.py"))


def load_core():
    """-> the reference's ``src.model.core`` module (ApertisConfig, ApertisLayer, ApertisForCausalLM ...)."""
    if not available():
        raise ImportError("the reference is not installed under baseline/_ref (run __graft_entry__.build() where "
                          "/root/reference exists)")
    src = sys.modules.get("src")
    if src is None or REF_ROOT not in list(getattr(src, "__path__", [])):
        src = types.ModuleType("src")
        src.__path__ = [REF_ROOT]
        sys.modules["src"] = src
    return importlib.import_module("src.model.core")
