"""`ChemeleonB200`: the reverse-diffusion sampler behind the reference's API.

Mirrors `Chemeleon.sample(text_input, n_atoms, n_samples, cond_scale=2.0,
step_lr=1e-5, return_trajectory=False, stream=False)` and the ragged
`_sample_generator(natoms, texts, cond_scale, step_lr)`
(chemeleon/modules/chemeleon.py:469-490, 305-467).  One timestep = one call of
`cb2_sampler_step` (film cond, predictor forward cond|null, predictor update,
corrector forward, corrector update), captured once in a CUDA graph and replayed
T times with a device-side timestep counter; nothing crosses the host/device
boundary inside the loop unless the caller asks for the per-step stream.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings
from typing import Callable, Iterator, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, schedules
from .atoms import state_to_atoms
from .config import SamplerConfig
from .engine import DecoderEngine, _stream_ptr
from .topology import BatchTopology

LATTICE_MASK = [[1, 0, 1], [1, 1, 1], [0, 0, 1]]  # chemeleon.py:70-72


class InjectedNoise:
    """Noise tensors in the reference's draw order (chemeleon.py:347-349,400-404,418,435,455).

    l_T [B,3,3], x_T [N,3]; rand_a [S,N,104], rand_l [S,B,3,3], rand_x / rand_x2 [S,N,3]
    with slice s belonging to timestep t_start - s (zeros where t == 1)."""

    def __init__(self, l_T, x_T, rand_a, rand_l, rand_x, rand_x2, t_start: int):
        self.l_T, self.x_T = l_T, x_T
        self.rand_a, self.rand_l, self.rand_x, self.rand_x2 = rand_a, rand_l, rand_x, rand_x2
        self.t_start = int(t_start)


class TextCondition:
    """What a batch is conditioned on: FiLM text rows [R,1024] (device) and, for every (variant, crystal)
    in the order [cond crystals | null crystals], the row it uses (int32 [2B])."""

    def __init__(self, rows: torch.Tensor, row_of: torch.Tensor):
        self.rows, self.row_of = rows, row_of


class SamplerRun:
    """Device buffers + (optionally captured) step for one batch on one GPU."""

    def __init__(self, engine: DecoderEngine, natoms: Sequence[int], cond: Optional["TextCondition"],
                 cond_scale: float, step_lr: float,
                 noise: Optional[InjectedNoise] = None, seed: int = 0,
                 graph_gid: Optional[Sequence[int]] = None, use_cuda_graph: bool = True):
        self.eng = engine
        cfg = engine.cfg
        dev = engine.device
        self.text_guide = cond is not None
        V = 2 if self.text_guide else 1
        self.topo: BatchTopology = engine.topology(natoms, V)
        B, N = self.topo.B, self.topo.N
        self.B, self.N, self.V = B, N, V
        self.T = cfg.timesteps
        with torch.cuda.device(dev):
            self.ws = engine.workspace(self.topo)
            self.a = torch.zeros(N, dtype=torch.int64, device=dev)
            self.x = torch.zeros(N, 3, dtype=torch.float32, device=dev)
            self.l = torch.zeros(B, 9, dtype=torch.float32, device=dev)
            self.t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self.flags = torch.zeros(max(B, 1), dtype=torch.int32, device=dev)
            coef = schedules.coefficient_table(cfg.timesteps, engine.w.sigmas_norm, step_lr, cfg.sigma_begin,
                                               cfg.sigma_end, cfg.beta_schedule, engine.w.q_mats,
                                               engine.w.q_one_step_mats)
            self.coef = coef.to(dev)
            # FiLM text rows [<= V*B + 1, 1024] + the row every (variant, crystal) uses: crystals with the
            # same prompt share a row, all unconditional rows are one row (k_film_cond gathers)
            self.text_part = torch.zeros(V * B + 1, 2 * cfg.hidden_dim, dtype=torch.float32, device=dev)
            self.text_row = torch.zeros(V * B, dtype=torch.int32, device=dev)
            if self.text_guide:
                self._set_condition(cond)
            else:
                self.text_part[0] = engine.w.film_b_cond
            gid = np.arange(B, dtype=np.int64) if graph_gid is None else np.asarray(graph_gid, dtype=np.int64)
            self.graph_gid = torch.from_numpy(gid).to(dev)
        self.noise = noise
        if noise is not None:
            self._noise_dev = [t.to(dev, torch.float32).contiguous() for t in
                               (noise.rand_a, noise.rand_l.reshape(noise.rand_l.shape[0], -1, 9), noise.rand_x,
                                noise.rand_x2)]
        self.state = _lib.State()
        self.state.atom_types, self.state.frac_coords = self.a.data_ptr(), self.x.data_ptr()
        self.state.lattices, self.state.t_dev, self.state.flags = (self.l.data_ptr(), self.t_dev.data_ptr(),
                                                                  self.flags.data_ptr())
        args = _lib.StepArgs()
        args.coef, args.text_part = self.coef.data_ptr(), self.text_part.data_ptr()
        args.text_row = self.text_row.data_ptr()
        args.cond_scale = float(cond_scale)
        args.timesteps = cfg.timesteps
        args.precision = engine.precision
        if noise is not None:
            args.noise_mode, args.t_start = 0, noise.t_start
            args.rand_a, args.rand_l, args.rand_x, args.rand_x2 = [t.data_ptr() for t in self._noise_dev]
        else:
            args.noise_mode, args.t_start = 1, cfg.timesteps
        args.seed = int(seed) & (2 ** 64 - 1)
        self.seed_dev = torch.tensor([int(seed) & (2 ** 63 - 1)], dtype=torch.int64, device=dev)
        args.seed_dev = self.seed_dev.data_ptr()
        args.graph_gid = self.graph_gid.data_ptr()
        self.args = args
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self.t_host = 0          # host mirror of the device timestep counter (guards against stepping past t = 1)
        self.busy = False        # owned by a live generator: must not be handed out again from the run cache

    # -- re-use of a captured run with new conditioning (pointers stay valid) -------
    def _set_condition(self, cond: "TextCondition") -> None:
        rows, row_of = cond.rows, cond.row_of
        if row_of.numel() != self.V * self.B or rows.shape[0] > self.text_part.shape[0]:
            raise ValueError("text condition does not match the batch (one row index per variant and crystal)")
        if int(row_of.max()) >= rows.shape[0] or int(row_of.min()) < 0:
            raise ValueError("text condition: row index out of range")
        self.text_part[: rows.shape[0]].copy_(rows, non_blocking=True)
        self.text_row.copy_(row_of.to(torch.int32), non_blocking=True)

    def reconfigure(self, cond: Optional["TextCondition"], seed: int,
                    graph_gid: Optional[Sequence[int]] = None) -> None:
        if self.text_guide:
            self._set_condition(cond)
        self.seed_dev.fill_(int(seed) & (2 ** 63 - 1))
        if graph_gid is not None:
            self.graph_gid.copy_(torch.as_tensor(np.asarray(graph_gid, dtype=np.int64)), non_blocking=True)

    # -- state ---------------------------------------------------------------
    def set_state(self, a: torch.Tensor, x: torch.Tensor, l: torch.Tensor, t: int) -> None:
        dev = self.eng.device
        self.a.copy_(a.to(dev, torch.int64))
        self.x.copy_(x.to(dev, torch.float32))
        self.l.copy_(l.to(dev, torch.float32).reshape(-1, 9))
        self.t_dev.fill_(int(t))
        self.t_host = int(t)
        self.flags.zero_()

    def init_state(self, l_T: torch.Tensor, x_T: torch.Tensor, t_start: Optional[int] = None) -> None:
        """a_T = 0, l_T masked, x_T wrapped (chemeleon.py:347-361)."""
        dev = self.eng.device
        mask = torch.tensor(LATTICE_MASK, dtype=torch.float32, device=dev)
        l = l_T.to(dev, torch.float32).reshape(-1, 3, 3) * mask
        x = torch.remainder(x_T.to(dev, torch.float32), 1.0)
        self.set_state(torch.zeros(self.N, dtype=torch.int64, device=dev), x, l,
                       self.T if t_start is None else t_start)

    def get_state(self):
        return self.a.clone(), self.x.clone(), self.l.clone().view(-1, 3, 3)

    # -- stepping --------------------------------------------------------------
    def _step_eager(self) -> None:
        _lib.check(self.eng.lib.cb2_sampler_step(C.byref(self.eng.model), self.topo.byref(), C.byref(self.state),
                                                 C.byref(self.args), self.ws.data_ptr(), self.ws.numel(),
                                                 _stream_ptr()), "cb2_sampler_step")

    def capture(self) -> None:
        """Warm up once (module load), restore the state, then capture one timestep."""
        if self._graph is not None or not self.use_cuda_graph:
            return
        saved = (self.a.clone(), self.x.clone(), self.l.clone(), self.t_dev.clone(), self.flags.clone())
        if int(saved[3].item()) < 1:
            self.t_dev.fill_(1)
        if self.noise is not None:   # keep the warm-up / capture steps inside the injected tensors
            self.t_dev.fill_(self.noise.t_start)
        s = torch.cuda.Stream(device=self.eng.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step_eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize(self.eng.device)
        for dst, src in zip((self.a, self.x, self.l, self.t_dev, self.flags), saved):
            dst.copy_(src)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            self._step_eager()
        for dst, src in zip((self.a, self.x, self.l, self.t_dev, self.flags), saved):
            dst.copy_(src)
        self._graph = g

    def step(self) -> None:
        if self.t_host < 1:
            raise RuntimeError("SamplerRun.step(): the run is finished (t < 1); call set_state/init_state first")
        if self.noise is not None and self.t_host > self.noise.t_start:
            raise RuntimeError("SamplerRun.step(): timestep above the first slice of the injected noise")
        if self.noise is not None and self.noise.t_start - self.t_host >= self._noise_dev[0].shape[0]:
            raise RuntimeError("SamplerRun.step(): injected noise tensors are exhausted")
        self.t_host -= 1
        if self.use_cuda_graph:
            if self._graph is None:
                self.capture()
            self._graph.replay()
        else:
            self._step_eager()

    def run(self, n_steps: int) -> None:
        for _ in range(n_steps):
            self.step()


class ChemeleonB200:
    """B200 sampler with the reference `Chemeleon` sampling API."""

    def __init__(self, source, cfg: Optional[SamplerConfig] = None, text_encoder=None, device="cuda",
                 precision: str = "fp32", use_cuda_graph: bool = True):
        """`source`: reference `Chemeleon` module, state_dict or checkpoint dict.
        `text_encoder`: object with the reference `TextEncoder.get_text_embeds(texts, cond_drop_prob, device)`
        interface (chemeleon/text_encoder/text_encoder.py:186-205); only needed for string prompts."""
        if cfg is None and hasattr(source, "hparams"):
            cfg = SamplerConfig.from_hparams(source.hparams)
        if text_encoder is None and hasattr(source, "text_encoder"):
            text_encoder = source.text_encoder
        self.cfg = cfg or SamplerConfig()
        self.engine = DecoderEngine(source, self.cfg, device, precision)
        self.device = self.engine.device
        self.text_encoder = text_encoder
        self.text_guide = self.cfg.text_guide
        self.use_cuda_graph = use_cuda_graph
        self.hparams = self.cfg
        self._run_cache = {}
        self.prompt_embeds = {}          # prompt -> language-model embedding [embed_dim] (host)
        self.source_hparams = dict(getattr(source, "hparams", {}) or {}) if not isinstance(source, dict) else \
            dict(source.get("hyper_parameters", {}))

    # -- loaders (reference: chemeleon.py:97-135) ----------------------------------
    @classmethod
    def load_from_checkpoint(cls, path_ckpt: str, text_encoder=None, **kw) -> "ChemeleonB200":
        if not os.path.exists(path_ckpt):
            raise FileNotFoundError(f"{path_ckpt}: checkpoints are not bundled and cannot be downloaded offline")
        ckpt = torch.load(path_ckpt, map_location="cpu", weights_only=False)
        cfg = SamplerConfig.from_hparams(ckpt.get("hyper_parameters", {}))
        model = cls(ckpt["state_dict"], cfg, text_encoder=text_encoder, **kw)
        model.source_hparams = dict(ckpt.get("hyper_parameters", {}))
        return model

    @classmethod
    def load_general_text_model(cls, checkpoint_dir: Optional[str] = None, **kw) -> "ChemeleonB200":
        d = checkpoint_dir or os.environ.get("CHEMELEON_CHECKPOINT_DIR", "checkpoints")
        return cls.load_from_checkpoint(os.path.join(d, "chemeleon-7fsg68c3.ckpt"), **kw)

    @classmethod
    def load_composition_model(cls, checkpoint_dir: Optional[str] = None, **kw) -> "ChemeleonB200":
        d = checkpoint_dir or os.environ.get("CHEMELEON_CHECKPOINT_DIR", "checkpoints")
        return cls.load_from_checkpoint(os.path.join(d, "chemeleon-fksq6cgp.ckpt"), **kw)

    def eval(self):
        return self

    # -- text ------------------------------------------------------------------
    def set_prompt_embedding(self, prompt: str, embedding: torch.Tensor) -> None:
        """Register the language-model embedding [embed_dim] of a prompt (what the reference's
        `TextEncoder.text_encode` returns: BERT class token, CrystalClip projection applied), so that
        `sample(text_input=prompt, ...)` runs without the language model at hand."""
        self.prompt_embeds[str(prompt)] = embedding.detach().reshape(-1).to(torch.float32).cpu()

    def condition_from_embeddings(self, B: int, text_embeds: torch.Tensor,
                                  null_text_embeds: torch.Tensor) -> TextCondition:
        """Projected text embeddings (outputs of `TextEncoder.get_text_embeds`): cond [B,512], null [1 or B,512]."""
        dev = self.device
        text = text_embeds.to(dev, torch.float32, non_blocking=True)
        null = null_text_embeds.to(dev, torch.float32, non_blocking=True)
        if text.shape[0] != B or null.shape[0] not in (1, B):
            raise ValueError("text embeddings must have one row per crystal (null: one row, or one per crystal)")
        rows = self.engine.text_part(torch.cat([text, null], dim=0))
        idx = torch.arange(B, dtype=torch.int32, device=dev)
        null_rows = idx + B if null.shape[0] == B else torch.full((B,), B, dtype=torch.int32, device=dev)
        return TextCondition(rows, torch.cat([idx, null_rows]))

    def condition_from_encoder(self, encoder_embeds: torch.Tensor, prompt_ids: Sequence[int]) -> TextCondition:
        """Language-model embeddings [P, embed_dim] of the DISTINCT prompts + the prompt of every crystal:
        the text tail (text_emb MLP, learned null embedding, FiLM fold) runs on the device, P + 1 rows."""
        rows = self.engine.text_condition(encoder_embeds)
        P = rows.shape[0] - 1
        ids = torch.as_tensor(list(prompt_ids), dtype=torch.int32)
        if ids.numel() and (int(ids.max()) >= P or int(ids.min()) < 0):
            raise ValueError("prompt id out of range")
        row_of = torch.cat([ids, torch.full_like(ids, P)]).to(self.device)
        return TextCondition(rows, row_of)

    def _embed_texts(self, texts: Sequence[str]):
        """Prompts -> conditioning.  With the checkpoint's own text tail (text_encoder.text_emb.*) only the
        language-model embedding of each DISTINCT prompt is needed: from `prompt_embeds`
        (`set_prompt_embedding`) or from `text_encoder.text_encode(prompts, device)`
        (text_encoder/text_encoder.py:129-184).  Otherwise `text_encoder.get_text_embeds` is called like the
        reference does (chemeleon.py:364-377)."""
        texts = [str(t) for t in texts]
        if self.engine.has_text_tail:
            distinct = list(dict.fromkeys(texts))
            missing = [p for p in distinct if p not in self.prompt_embeds]
            if missing and self.text_encoder is not None and hasattr(self.text_encoder, "text_encode"):
                enc = self.text_encoder.text_encode(missing, self.device).detach().to(torch.float32).cpu()
                for p, e in zip(missing, enc):
                    self.prompt_embeds[p] = e
                missing = []
            if not missing:
                enc = torch.stack([self.prompt_embeds[p] for p in distinct])
                index = {p: i for i, p in enumerate(distinct)}
                return self.condition_from_encoder(enc, [index[t] for t in texts]), None
        if self.text_encoder is None or not hasattr(self.text_encoder, "get_text_embeds"):
            raise ValueError(
                "string prompts need either the language-model embedding of every prompt (set_prompt_embedding / a "
                "text_encoder with text_encode) together with a checkpoint that holds text_encoder.text_emb.*, or a "
                "text_encoder with get_text_embeds; alternatively pass text_embeds / null_text_embeds")
        te = self.text_encoder.get_text_embeds(list(texts), cond_drop_prob=0.0, device=self.device)
        ne = self.text_encoder.get_text_embeds(list(texts), cond_drop_prob=1.0, device=self.device)
        return te.detach(), ne.detach()

    # -- core ------------------------------------------------------------------
    def make_run(self, natoms: Sequence[int], text_embeds=None, null_text_embeds=None, cond_scale: float = 2.0,
                 step_lr: float = 1e-5, noise: Optional[InjectedNoise] = None, seed: int = 0,
                 graph_gid=None) -> SamplerRun:
        if self.text_guide and text_embeds is None:
            raise ValueError("text_guide model: text embeddings are required")
        cond = None
        if self.text_guide:
            cond = text_embeds if isinstance(text_embeds, TextCondition) else \
                self.condition_from_embeddings(len(natoms), text_embeds, null_text_embeds)
        if noise is not None:  # parity mode: injected tensors are baked into the step arguments
            return SamplerRun(self.engine, natoms, cond, cond_scale, step_lr, noise, seed, graph_gid,
                              self.use_cuda_graph)
        # production mode: a captured run is re-used for every later call with the same batch
        # shape (e.g. composition sweeps); only the conditioning, seed and sample ids change.
        key = (tuple(int(n) for n in natoms), float(cond_scale), float(step_lr))
        run = self._run_cache.get(key)
        if run is not None and run.busy:
            # a live generator (stream=True) owns the cached run: its state must not be shared
            return SamplerRun(self.engine, natoms, cond, cond_scale, step_lr, None, seed, graph_gid,
                              self.use_cuda_graph)
        if run is None:
            if len(self._run_cache) >= 4:
                self._run_cache = {k: r for k, r in self._run_cache.items() if r.busy}
            run = SamplerRun(self.engine, natoms, cond, cond_scale, step_lr, None, seed, graph_gid,
                             self.use_cuda_graph)
            self._run_cache[key] = run
        else:
            gid = graph_gid if graph_gid is not None else np.arange(run.B, dtype=np.int64)
            run.reconfigure(cond, seed, gid)
        return run

    def initial_noise(self, B: int, N: int, seed: int):
        """l_T, x_T from a device generator (production mode; global order => sharding invariant)."""
        g = torch.Generator(device=self.device).manual_seed(int(seed))
        l_T = torch.randn(B, 3, 3, generator=g, device=self.device)
        x_T = torch.randn(N, 3, generator=g, device=self.device)
        return l_T, x_T

    @torch.no_grad()
    def sample_states(self, natoms: Sequence[int], text_embeds=None, null_text_embeds=None,
                      cond_scale: float = 2.0, step_lr: float = 1e-5, noise: Optional[InjectedNoise] = None,
                      seed: int = 0, t_stop: int = 0, init_state=None, t_start: Optional[int] = None,
                      per_step: Optional[Callable] = None, graph_gid=None, init_noise=None):
        """Run t = t_start .. t_stop+1 and return the device state (a, x, l) at t_stop."""
        natoms = [int(n) for n in natoms]
        with torch.cuda.device(self.device):
            run = self.make_run(natoms, text_embeds, null_text_embeds, cond_scale, step_lr, noise, seed, graph_gid)
            t0 = self.cfg.timesteps if t_start is None else int(t_start)
            if init_state is not None:
                run.set_state(*init_state, t0)
            else:
                if noise is not None:
                    l_T, x_T = noise.l_T, noise.x_T
                elif init_noise is not None:
                    l_T, x_T = init_noise
                else:
                    l_T, x_T = self.initial_noise(run.B, run.N, seed)
                run.init_state(l_T, x_T, t0)
            for t in range(t0, t_stop, -1):
                run.step()
                if per_step is not None:
                    per_step(t - 1, run)
            a, x, l = run.get_state()
            self.last_flags = run.flags.clone()
        return a, x, l

    def check_flags(self, flags: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Guard bits of the last run (host int32 [B]); warns when any crystal is flagged.
        Bit 1 (`_lib.FLAG_NONFINITE`): the update produced NaN/Inf; bit 2 (`_lib.FLAG_TC_RANGE`):
        tensor-core mode met a cell outside its fp16 range -- re-run those samples with precision="fp32"."""
        f = (self.last_flags if flags is None else flags).cpu()
        bad = int((f != 0).sum())
        if bad:
            nf = int(((f & _lib.FLAG_NONFINITE) != 0).sum())
            rg = int(((f & _lib.FLAG_TC_RANGE) != 0).sum())
            warnings.warn(f"chemeleon_b200: {bad} of {f.numel()} structures are flagged "
                          f"({nf} non-finite, {rg} outside the tensor-core range); see ChemeleonB200.last_flags",
                          RuntimeWarning, stacklevel=2)
        return f

    def _to_atoms(self, a, x, l, natoms):
        return state_to_atoms(a.cpu().numpy(), x.cpu().numpy(), l.reshape(-1, 9).cpu().numpy(), natoms)

    # -- reference API ---------------------------------------------------------------
    def _sample_generator(self, natoms: Union[int, List[int]], texts: Optional[Union[str, List[str]]] = None,
                          cond_scale: float = 2.0, step_lr: float = 1e-5, *, text_embeds=None,
                          null_text_embeds=None, noise=None, seed: int = 0, frames: bool = False,
                          depth: int = 4) -> Iterator[List]:
        """Yields once per timestep (T items), like the reference generator: `List[Atoms]`, or with
        `frames=True` the compact wire-format `streaming.Frame` of the step.  The GPU is not stalled:
        the sampling stream runs up to `depth` timesteps ahead of the frame being handed out (packed
        on the device, copied into a pinned ring buffer on a side stream)."""
        from .streaming import FrameStreamer

        if isinstance(natoms, int):
            natoms = [natoms]
        if isinstance(texts, str):
            texts = [texts]
        if texts is not None and len(texts) != len(natoms):
            raise ValueError("natoms and texts must have the same number of elements.")
        if self.text_guide and text_embeds is None:
            text_embeds, null_text_embeds = self._embed_texts(texts)
        natoms = [int(n) for n in natoms]
        with torch.cuda.device(self.device):
            run = self.make_run(natoms, text_embeds, null_text_embeds, cond_scale, step_lr, noise, seed)
            run.busy = True
            try:
                if noise is not None:
                    run.init_state(noise.l_T, noise.x_T)
                else:
                    run.init_state(*self.initial_noise(run.B, run.N, seed))
                streamer = FrameStreamer(run, depth)
                T = self.cfg.timesteps
                done = 0
                for _ in range(T):
                    while done < T and streamer.pending < streamer.depth:
                        with torch.cuda.device(self.device):
                            run.step()
                            streamer.push()
                        done += 1
                    frame = streamer.pop()
                    yield frame if frames else frame.to_atoms(natoms)
                self.last_flags = run.flags.clone()
            finally:
                run.busy = False

    def sample(self, text_input: str, n_atoms: int, n_samples: int, cond_scale: float = 2.0,
               step_lr: float = 1e-5, return_trajectory: bool = False, stream: bool = False, **kw):
        natoms = [n_atoms] * n_samples
        texts = [text_input] * n_samples
        if stream:
            return self._sample_generator(natoms, texts, cond_scale, step_lr, **kw)
        if return_trajectory:
            return list(self._sample_generator(natoms, texts, cond_scale, step_lr, **kw))
        return self.sample_batch(natoms, texts, cond_scale=cond_scale, step_lr=step_lr, **kw)

    def sample_batch(self, natoms: Sequence[int], texts: Optional[Sequence[str]] = None, *, text_embeds=None,
                     null_text_embeds=None, cond_scale: float = 2.0, step_lr: float = 1e-5, noise=None,
                     seed: int = 0, t_stop: int = 0, return_flags: bool = False):
        """Ragged entry point (the use-case of the reference's stale list-based callers,
        scripts/evaluate.py:97-99): final structures only, one D2H at the end.  The per-crystal guard
        bits are checked (RuntimeWarning) and returned with `return_flags=True`."""
        if self.text_guide and text_embeds is None:
            text_embeds, null_text_embeds = self._embed_texts(texts)
        a, x, l = self.sample_states(natoms, text_embeds, null_text_embeds, cond_scale, step_lr, noise, seed,
                                     t_stop=t_stop)
        flags = self.check_flags()
        atoms = self._to_atoms(a, x, l, [int(n) for n in natoms])
        return (atoms, flags) if return_flags else atoms

    def sample_batch_valid(self, natoms: Sequence[int], texts: Optional[Sequence[str]] = None, *,
                           target_composition: Optional[str] = None, max_length: float = 60.0,
                           min_distance: float = 0.5, **kw):
        """`sample_batch` followed by the reference's validity filters, evaluated on the device
        (scripts/evaluate.py:177-189, sample_target_composition.py:57-62; see validity.py).
        Returns (atoms of the structures that pass, flags int32[B] of all structures)."""
        from .validity import validity_flags

        natoms = [int(n) for n in natoms]
        text_embeds, null_text_embeds = kw.pop("text_embeds", None), kw.pop("null_text_embeds", None)
        if self.text_guide and text_embeds is None:
            text_embeds, null_text_embeds = self._embed_texts(texts)
        a, x, l = self.sample_states(natoms, text_embeds, null_text_embeds, **kw)
        self.check_flags()
        with torch.cuda.device(self.device):
            flags, _, _ = validity_flags(a, x, l, natoms, target_composition, max_length, min_distance)
        flags = flags.cpu()
        atoms = self._to_atoms(a, x, l, natoms)
        return [at for at, f in zip(atoms, flags.tolist()) if f == 0], flags
