"""Named model configurations of the reference's shipped YAMLs, and a Hydra-free ``instantiate``.

The reference selects its model classes through Hydra ``_target_`` strings
(``hydra.utils.instantiate(cfg.model)``, reference experiments/base_experiment.py:116).  Hydra is not a
dependency of this package: :func:`instantiate` resolves the same recursive ``_target_`` mappings (as
loaded by ``yaml.safe_load`` from e.g. reference configs/model/cfm/cfm_ds2_electrons.yaml), and
:data:`TARGETS` maps the reference's class paths onto their B200 drop-ins, so a reference YAML can be used
unchanged (``instantiate(cfg, remap=True)``) or with only its ``_target_`` strings edited.

:data:`MODELS` restates the hyper-parameters of the shipped shape-model configs (file cited per entry) for
hosts that do not read YAML (bench.py, tests).
"""
from __future__ import annotations

import importlib
from typing import Any, Dict, Mapping

__all__ = ["MODELS", "TARGETS", "build", "instantiate"]

# reference class path -> drop-in
TARGETS = {
    "nn.vit.ViT": "vit4hep_b200.ViT",
    "models.base_model.CFM": "vit4hep_b200.CFM",
    "experiments.calochallenge.calochallenge_cfm.model.CaloChallengeCFM": "vit4hep_b200.CaloChallengeCFM",
    "experiments.calochallenge.calochallenge_cfm.model.CaloChallengeCFM_DS1": "vit4hep_b200.CaloChallengeCFM_DS1",
    "experiments.calogan.model.CaloGANCFM": "vit4hep_b200.CaloGANCFM",
    "experiments.calohadronic.model.CaloHadCFM": "vit4hep_b200.CaloHadCFM",
    "experiments.lemurs.model.LEMURSCFM": "vit4hep_b200.LEMURSCFM",
    "nn.cfm.transformer_cfm.ParallelTransformer": "vit4hep_b200.ParallelTransformer",
}


def _locate(path: str):
    module, _, name = path.rpartition(".")
    if not module:
        raise ImportError(f"_target_ {path!r} is not a dotted path")
    return getattr(importlib.import_module(module), name)


def instantiate(cfg: Any, remap: bool = False, **overrides):
    """Recursive ``_target_`` instantiation with Hydra's semantics for the subset the reference uses: a mapping
    with ``_target_`` becomes ``target(**other_items)`` after its values were instantiated; mappings without
    ``_target_`` and lists are walked; everything else is passed through.  ``remap`` replaces the reference's
    class paths by their drop-ins (:data:`TARGETS`); ``overrides`` update the top-level keyword arguments."""
    if isinstance(cfg, Mapping):
        if "_target_" in cfg:
            target = cfg["_target_"]
            if remap:
                target = TARGETS.get(target, target)
            kwargs = {k: instantiate(v, remap) for k, v in cfg.items() if k != "_target_"}
            kwargs.update(overrides)
            return _locate(target)(**kwargs)
        return {k: instantiate(v, remap) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return [instantiate(v, remap) for v in cfg]
    return cfg


def _vit(patch_dim, num_patches, condition_dim):
    return dict(dim=3, condition_dim=condition_dim, hidden_dim=480, out_channels=1, depth=6, num_heads=6,
                mlp_ratio=4, attn_drop=0.0, proj_drop=0.0, pos_embedding_coords="cylindrical", temperature=10000,
                learn_pos_embed=True, causal_attn=False, checkpoint_grads=False, num_patches=num_patches,
                patch_dim=patch_dim, use_torch_sdpa=False)


_ODE = dict(method="rk4", options=dict(step_size=0.05))
_DS1_PHOTONS = [[1, 8, 5], [1, 16, 10], [1, 19, 10], [1, 5, 5], [1, 5, 5]]
_DS1_PIONS = [[1, 8, 5], [1, 10, 10], [1, 10, 10], [1, 5, 5], [1, 15, 10], [1, 16, 10], [1, 10, 5]]


def _vol(shapes):
    return [s[0] * s[1] * s[2] for s in shapes]


# name -> wrapper config in the reference's YAML structure (``_target_`` already pointing at the drop-ins)
MODELS: Dict[str, dict] = {
    # reference configs/model/cfm/cfm_ds2_electrons.yaml
    "ds2": dict(_target_="vit4hep_b200.CaloChallengeCFM", in_channels=1, shape=[45, 16, 9], patch_shape=[3, 16, 1],
                time_distribution="uniform", trajectory="linear", odeint_kwargs=_ODE,
                net=dict(_target_="vit4hep_b200.ViT", param=_vit(48, [[15, 1, 9]], 46))),
    # reference configs/model/cfm/cfm_ds3_electrons.yaml
    "ds3": dict(_target_="vit4hep_b200.CaloChallengeCFM", in_channels=1, shape=[45, 50, 18], patch_shape=[3, 10, 3],
                time_distribution="uniform", trajectory="linear", odeint_kwargs=_ODE,
                net=dict(_target_="vit4hep_b200.ViT", param=_vit(90, [[15, 5, 6]], 46))),
    # reference configs/model/cfm_lemurs/cfm_lemurs.yaml
    "lemurs": dict(_target_="vit4hep_b200.LEMURSCFM", in_channels=1, shape=[45, 16, 9], patch_shape=[3, 16, 1],
                   time_distribution="uniform", trajectory="linear", odeint_kwargs=_ODE,
                   net=dict(_target_="vit4hep_b200.ViT", param=_vit(48, [[15, 1, 9]], 53))),
    # reference configs/model/cfm/cfm_ds1_photons.yaml
    "ds1_photons": dict(_target_="vit4hep_b200.CaloChallengeCFM_DS1", in_channels=1, shape=[368 + 72],
                        list_shape=_DS1_PHOTONS, list_edges=_vol(_DS1_PHOTONS), patch_shape=[1, 1, 5],
                        time_distribution="uniform", trajectory="linear", odeint_kwargs=_ODE,
                        net=dict(_target_="vit4hep_b200.ViT",
                                 param=_vit(5, [[1, 8, 1], [1, 16, 2], [1, 19, 2], [1, 5, 1], [1, 5, 1]], 6))),
    # reference configs/model/cfm/cfm_ds1_pions.yaml
    "ds1_pions": dict(_target_="vit4hep_b200.CaloChallengeCFM_DS1", in_channels=1, shape=[625],
                      list_shape=_DS1_PIONS, list_edges=_vol(_DS1_PIONS), patch_shape=[1, 1, 5],
                      time_distribution="uniform", trajectory="linear", odeint_kwargs=_ODE,
                      net=dict(_target_="vit4hep_b200.ViT",
                               param=_vit(5, [[1, 8, 1], [1, 10, 2], [1, 10, 2], [1, 5, 1], [1, 15, 2], [1, 16, 2],
                                              [1, 10, 1]], 8))),
    # reference configs/model/cfm_calogan/cfm_eplus.yaml
    "calogan": dict(_target_="vit4hep_b200.CaloGANCFM", in_channels=1, shape=[504],
                    list_shape=[[1, 96, 3], [1, 12, 12], [1, 6, 12]], list_edges=[288, 144, 72],
                    list_patch_shape=[[1, 6, 1], [1, 2, 3], [1, 2, 3]],
                    time_distribution="uniform", trajectory="linear", odeint_kwargs=_ODE,
                    net=dict(_target_="vit4hep_b200.ViT", param=_vit(6, [[1, 16, 3], [1, 6, 4], [1, 3, 4]], 4))),
    # reference configs/model/cfm_calohad/cfm_calohad.yaml
    "calohad": dict(_target_="vit4hep_b200.CaloHadCFM", in_channels=1, shape=[45450],
                    list_shape=[[10, 15, 15], [48, 30, 30]], list_edges=[2250, 43200],
                    list_patch_shape=[[5, 5, 3], [3, 5, 5]],
                    time_distribution="uniform", trajectory="linear", odeint_kwargs=_ODE,
                    net=dict(_target_="vit4hep_b200.ViT", param=_vit(75, [[2, 3, 5], [16, 6, 6]], 59))),
}


_ENERGY = dict(dims_in=45, dims_c=1, dim_embedding=64, nhead=4, num_encoder_layers=4, num_decoder_layers=4,
               dim_feedforward=512, dropout=0.0, activation="relu", embeds=True, encode_t_scale=30)
# reference configs/model/cfm/cfm_ds2_energy.yaml (cfm_ds3_energy.yaml is identical): the energy-ratio CFM that is
# sampled before the shape model
MODELS["ds2_energy"] = dict(_target_="vit4hep_b200.CFM", shape=[45], time_distribution="uniform", trajectory="linear",
                            odeint_kwargs=_ODE, net=dict(_target_="vit4hep_b200.ParallelTransformer", param=dict(_ENERGY)))
MODELS["ds3_energy"] = MODELS["ds2_energy"]


def build(name: str, precision: str = "bf16", **param_overrides):
    """The named model (wrapper + net) with the shipped hyper-parameters; ``param_overrides`` update the ViT's
    ``param`` mapping (e.g. hidden_dim=96, depth=2 for a small instance)."""
    import copy
    cfg = copy.deepcopy(MODELS[name])
    cfg["net"]["param"].update(param_overrides)
    cfg["net"]["param"]["precision"] = precision
    return instantiate(cfg)
