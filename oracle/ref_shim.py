"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference from /root/reference.

The reference (`ryannduma/chemeleon`) cannot be imported as-is in this image:
`chemeleon/__init__.py:1` pulls in torch_geometric, pytorch_lightning,
torchmetrics, ase, wandb and a BERT download.  This module registers small stub
modules for those third-party imports (they only supply index generation, a
module base class and the output container on the sampling path -- no floating
point math) and then imports the reference's own files from
where they lie.  Nothing from the reference is copied.

Only `tests/`, `tests/golden/make_golden.py` and `oracle/` self-checks may use
this.  It is unavailable on the GPU box (`/root/reference` does not travel);
callers must check `reference_available()`.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("CHEMELEON_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "chemeleon", "modules"))


class _FakeAtoms:
    """Minimal stand-in for ase.Atoms (numbers / cell / scaled positions / pbc)."""

    def __init__(self, numbers=None, cell=None, pbc=True):
        import numpy as np

        self.numbers = np.asarray(numbers)
        self.cell = np.asarray(cell, dtype=float)
        self.pbc = pbc
        self.scaled_positions = None

    def set_scaled_positions(self, pos):
        import numpy as np

        self.scaled_positions = np.asarray(pos, dtype=float)

    def __len__(self):
        return len(self.numbers)


def _fake_sort(atoms):
    # The tensor-level tests never depend on the ase sort; keep order.
    return atoms


def _install_stubs() -> None:
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tg_utils = types.ModuleType("torch_geometric.utils")
        tg_data = types.ModuleType("torch_geometric.data")

        def dense_to_sparse(adj):
            idx = adj.nonzero().t().contiguous()
            return idx, adj[idx[0], idx[1]]

        class Data:
            def __init__(self, **kw):
                self.__dict__.update(kw)

        class Batch:
            @staticmethod
            def from_data_list(data_list):
                b = Batch()
                natoms = torch.tensor([int(d.natoms) for d in data_list], dtype=torch.long)
                b.natoms = natoms
                b.num_graphs = len(data_list)
                b.num_nodes = int(natoms.sum())
                b.batch = torch.arange(len(data_list)).repeat_interleave(natoms)
                return b

            def to(self, device):
                self.natoms = self.natoms.to(device)
                self.batch = self.batch.to(device)
                return self

        tg_utils.dense_to_sparse = dense_to_sparse
        tg_data.Data = Data
        tg_data.Batch = Batch
        tg.utils = tg_utils
        tg.data = tg_data
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.utils"] = tg_utils
        sys.modules["torch_geometric.data"] = tg_data

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                if a and isinstance(a[0], dict):
                    self.hparams = types.SimpleNamespace(**a[0])

            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def log(self, *a, **k):
                pass

        pl.LightningModule = LightningModule
        sys.modules["pytorch_lightning"] = pl

    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")

        class MeanAbsoluteError(nn.Module):
            def forward(self, a, b):
                return (a - b).abs().mean()

        tm.MeanAbsoluteError = MeanAbsoluteError
        sys.modules["torchmetrics"] = tm

    try:
        import ase  # noqa: F401
    except Exception:
        ase = types.ModuleType("ase")
        ase.Atoms = _FakeAtoms
        ase_build = types.ModuleType("ase.build")
        ase_tools = types.ModuleType("ase.build.tools")
        ase_tools.sort = _fake_sort
        ase_build.tools = ase_tools
        ase.build = ase_build
        sys.modules["ase"] = ase
        sys.modules["ase.build"] = ase_build
        sys.modules["ase.build.tools"] = ase_tools

    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            sys.modules["wandb"] = types.ModuleType("wandb")


class FakeTextEncoder(nn.Module):
    """Replaces chemeleon.text_encoder.text_encoder.TextEncoder (BERT weights are
    not downloadable offline).  Returns fixed [B, text_dim] embeddings that the
    test sets: `cond` for cond_drop_prob=0 and `null` for cond_drop_prob=1, which
    is exactly how `_sample_generator` queries it (chemeleon.py:364-377)."""

    def __init__(self, *a, text_dim=512, **k):
        super().__init__()
        self.text_dim = text_dim
        self.cond = None  # [B, text_dim]
        self.null = None  # [1, text_dim] or [B, text_dim]

    def get_text_embeds(self, batch_text, cond_drop_prob, device):
        bsz = len(batch_text)
        if cond_drop_prob >= 1.0:
            e = self.null
        else:
            e = self.cond
        if e.shape[0] == 1:
            e = e.expand(bsz, -1)
        assert e.shape[0] == bsz
        return e.to(device).clone()


_REF = None


def load_reference():
    """Returns a namespace with the reference's own modules (imported unmodified)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_stubs()
    # namespace package so chemeleon/__init__.py (which imports everything) is skipped
    if "chemeleon" not in sys.modules or not hasattr(sys.modules["chemeleon"], "__path__") \
            or REFERENCE_ROOT not in str(sys.modules["chemeleon"].__path__):
        pkg = types.ModuleType("chemeleon")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "chemeleon")]
        sys.modules["chemeleon"] = pkg
        for sub in ("modules", "utils", "text_encoder"):
            m = types.ModuleType(f"chemeleon.{sub}")
            m.__path__ = [os.path.join(REFERENCE_ROOT, "chemeleon", sub)]
            sys.modules[f"chemeleon.{sub}"] = m
        # text encoder package: stub the heavy modules
        te = types.ModuleType("chemeleon.text_encoder.text_encoder")
        te.TextEncoder = FakeTextEncoder
        sys.modules["chemeleon.text_encoder.text_encoder"] = te
        cc = types.ModuleType("chemeleon.text_encoder.crystal_clip")

        class CrystalClip(nn.Module):
            @classmethod
            def load_from_checkpoint(cls, *a, **k):
                raise RuntimeError("checkpoints are not available offline")

        cc.CrystalClip = CrystalClip
        sys.modules["chemeleon.text_encoder.crystal_clip"] = cc
        dl = types.ModuleType("chemeleon.utils.download")
        dl.download_file = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("offline"))
        sys.modules["chemeleon.utils.download"] = dl

    ns = types.SimpleNamespace()
    ns.cspnet = importlib.import_module("chemeleon.modules.cspnet")
    ns.diff_utils = importlib.import_module("chemeleon.utils.diff_utils")
    ns.scatter = importlib.import_module("chemeleon.utils.scatter")
    ns.schema = importlib.import_module("chemeleon.modules.schema")
    ns.chemeleon = importlib.import_module("chemeleon.modules.chemeleon")
    _REF = ns
    return ns


def default_config() -> dict:
    """The reference's default hyper-parameters (chemeleon/config.py:28-42,45-61,64-69,95)."""
    return dict(
        hidden_dim=512, time_dim=128, text_dim=512, max_atoms=104, num_layers=6,
        act_fn="silu", dis_emb="sin", num_freqs=128, edge_style="fc", max_neighbors=20,
        cutoff=6.0, ln=True, ip=True, smooth=False, pred_atom_types=True,
        text_guide=True, text_targets=["composition"], trainable_text_encoder=False,
        text_encoder="lfoppiano/MatTPUSciBERT", text_embed_dim=768, max_text_len=256,
        cond_drop_prob=0.2, beta_schedule="cosine", timesteps=1000, max_num_atoms=50,
        cost_atom_types=1.0, cost_lattice=1.0, cost_coords=1.0, d3pm_hybrid_coeff=1.0,
        optimizer="adam", lr=1e-3, weight_decay=0, scheduler="reduce_on_plateau",
        patience=200, cond_scale=2.0,
    )


def build_reference_model(seed: int = 0, cfg: dict | None = None):
    """Random-init reference `Chemeleon` (checkpoints are absent offline)."""
    ref = load_reference()
    cfg = dict(default_config(), **(cfg or {}))
    torch.manual_seed(seed)
    model = ref.chemeleon.Chemeleon(cfg)
    model.eval()
    return model


# --------------------------------------------------------------------------
# the REAL reference TextEncoder with a stand-in language model
# --------------------------------------------------------------------------
class FakeLanguageModel:
    """Stands in for the BERT + CrystalClip pair (weights are not downloadable offline): a
    deterministic "class token" embedding per prompt.  Everything AFTER the language model --
    `text_emb`, `null_text_embeds`, the keep mask of `get_text_embeds` -- is the reference's own code.
    Passed to the reference `TextEncoder` as `pretrained_clip_model` (text_encoder.py:48-51)."""

    def __init__(self, table, dim):
        outer = self
        self.table, self.dim = table, dim          # prompt -> tensor [dim]
        self._batch = None

        class _Tokenizer:
            # tokenizer interface used by TextEncoder.text_encode (text_encoder.py:130-136)
            def batch_encode_plus(self, batch_text, **kw):
                outer._batch = list(batch_text)
                ids = torch.zeros(len(batch_text), 1, dtype=torch.long)
                return types.SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids))

        class _Encoder(nn.Module):
            # language-model interface (text_encoder.py:170-175): class token = last_hidden_state[:, 0]
            def forward(self, input_ids=None, attention_mask=None):
                emb = torch.stack([outer.table[p] for p in outer._batch]).to(input_ids.device)
                return types.SimpleNamespace(last_hidden_state=emb[:, None, :])

        self.tokenizer = _Tokenizer()
        self.text_encoder = _Encoder()
        self.text_proj = nn.Identity()


def load_reference_text_encoder():
    """The reference's real `TextEncoder` class, imported from its own file (the package-level
    `chemeleon.text_encoder.text_encoder` entry stays the light stub the sampler tests use)."""
    load_reference()
    import importlib.util

    pkg = sys.modules["chemeleon.text_encoder"]
    init = importlib.util.spec_from_file_location("_cb2_ref_te_init",
                                                  os.path.join(REFERENCE_ROOT, "chemeleon", "text_encoder", "__init__.py"))
    m = importlib.util.module_from_spec(init)
    init.loader.exec_module(m)
    pkg.MODEL_NAMES, pkg.ARTIFACT_PATHS = m.MODEL_NAMES, m.ARTIFACT_PATHS
    spec = importlib.util.spec_from_file_location(
        "_cb2_ref_text_encoder", os.path.join(REFERENCE_ROOT, "chemeleon", "text_encoder", "text_encoder.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.TextEncoder
