"""The reference's plugin hook without MONAI: `define_instance(args, key)`
(3d_ldm/utils.py:243-246) resolves MONAI-bundle syntax in the JSON config
(3d_ldm/config/config_train_32g.json:7-28,40-49):

    "_target_": dotted path to instantiate      "@name": reference to another config entry
    "$expr":    python expression ("$@image_channels" -> value of image_channels)

`monai.networks.nets.{AutoencoderKL,DiffusionModelUNet}`, `networks.*` (config_train_16g.json:8,40)
and `monai.networks.schedulers.*` / `monai.inferers.*` targets are mapped to the b200ldm classes,
so the reference's configs work unmodified.
"""
from __future__ import annotations

import importlib
import re
from typing import Any

_ALIASES = {
    "monai.networks.nets.AutoencoderKL": "b200ldm.nets.AutoencoderKL",
    "monai.networks.nets.DiffusionModelUNet": "b200ldm.nets.DiffusionModelUNet",
    "networks.AutoencoderKL": "b200ldm.nets.AutoencoderKL",
    "networks.DiffusionModelUNet": "b200ldm.nets.DiffusionModelUNet",
    "generative.networks.nets.AutoencoderKL": "b200ldm.nets.AutoencoderKL",
    "generative.networks.nets.DiffusionModelUNet": "b200ldm.nets.DiffusionModelUNet",
    "monai.networks.schedulers.DDPMScheduler": "b200ldm.schedulers.DDPMScheduler",
    "monai.networks.schedulers.DDIMScheduler": "b200ldm.schedulers.DDIMScheduler",
    "monai.inferers.LatentDiffusionInferer": "b200ldm.inferer.LatentDiffusionInferer",
    "monai.inferers.DiffusionInferer": "b200ldm.inferer.DiffusionInferer",
}
_REF = re.compile(r"@(\w+(?:[#.]\w+)*)")


def _locate(path: str):
    path = _ALIASES.get(path, path)
    mod, _, name = path.rpartition(".")
    if not mod:
        raise ValueError(f"_target_ {path!r} is not a dotted path")
    return getattr(importlib.import_module(mod), name)


class ConfigParser:
    """Minimal re-implementation of the subset of monai.bundle.ConfigParser the reference uses."""

    def __init__(self, config: dict):
        self.config = dict(config)
        self._resolving: set[str] = set()

    def _lookup(self, ref: str) -> Any:
        cur: Any = self.config
        for part in re.split(r"[#.]", ref):
            cur = cur[int(part)] if isinstance(cur, (list, tuple)) else cur[part]
        return cur

    def _resolve_ref(self, ref: str, instantiate: bool) -> Any:
        if ref in self._resolving:
            raise ValueError(f"circular reference @{ref}")
        self._resolving.add(ref)
        try:
            return self._resolve(self._lookup(ref), instantiate)
        finally:
            self._resolving.discard(ref)

    def _resolve(self, node: Any, instantiate: bool) -> Any:
        if isinstance(node, dict):
            if "_target_" in node:
                if node.get("_disabled_", False):
                    return None
                kwargs = {k: self._resolve(v, instantiate) for k, v in node.items() if not k.startswith("_")}
                if not instantiate:
                    return kwargs
                return _locate(node["_target_"])(**kwargs)
            return {k: self._resolve(v, instantiate) for k, v in node.items()}
        if isinstance(node, (list, tuple)):
            return [self._resolve(v, instantiate) for v in node]
        if isinstance(node, str):
            if node.startswith("$"):
                names: dict[str, Any] = {}

                def sub(m):
                    var = "__ref_" + re.sub(r"\W", "_", m.group(1))
                    names[var] = self._resolve_ref(m.group(1), instantiate)
                    return var

                return eval(_REF.sub(sub, node[1:]), {"__builtins__": {}}, names)  # noqa: S307 - config DSL
            m = _REF.fullmatch(node)
            if m:
                return self._resolve_ref(m.group(1), instantiate)
        return node

    def parse(self, reset: bool = True) -> None:  # API compatibility
        return None

    def get_parsed_content(self, id: str, instantiate: bool = True) -> Any:
        return self._resolve(self._lookup(id), instantiate)


def define_instance(args, instance_def_key: str):
    """Drop-in for 3d_ldm/utils.py:243-246."""
    parser = ConfigParser(vars(args) if not isinstance(args, dict) else args)
    parser.parse(True)
    return parser.get_parsed_content(instance_def_key, instantiate=True)
