"""Drop-in `PerceiverEncoder` / `PerceiverDecoder` (reference: perceiver_io/perceiver.py:13-107, :110-180).

Same constructor kwargs, attribute / parameter names and forward signatures; this is where the hot path is
orchestrated (cross-attend -> num_blocks x shared self-attends; decoder cross-attend -> final_layer) and where the
multi-GPU partitioning of SURVEY.md §8(e) hooks in (`parallel.py`).
"""
from __future__ import annotations


import torch
import torch.nn as nn

from . import engine, ops, validate
from .inputs import PositionedInput
from .primitives import CrossAttention, SelfAttention, lecun_normal_, make_cross_attention_mask  # noqa: F401


class TrainablePositionEncoding(nn.Module):
    """The latent array holder: parameter `pos_embs` [index_dim, num_channels]
    (reference: position_encoding.py:104-124, used at perceiver.py:64-67)."""

    def __init__(self, index_dim, num_channels: int = 128, init_scale: float = 0.02):
        super().__init__()
        self.pos_embs = nn.Parameter(torch.zeros((index_dim, num_channels)))
        nn.init.trunc_normal_(self.pos_embs, std=init_scale)
        self._output_channels = num_channels

    def forward(self, batch_size, pos=None):
        del pos
        pos_embs = self.pos_embs
        if batch_size is not None:
            pos_embs = torch.broadcast_to(self.pos_embs[None, :, :], (batch_size,) + self.pos_embs.shape)
        return pos_embs

    def n_output_channels(self):
        return self._output_channels


class PerceiverEncoder(nn.Module):
    """The Perceiver encoder (reference: perceiver.py:13-107)."""

    def __init__(self, num_input_channels: int, num_self_attends_per_block: int = 6, num_blocks: int = 8,
                 num_latents: int = 512, num_latent_channels: int = 1024, qk_channels: int = None,
                 v_channels: int = None, num_cross_attend_heads: int = 1, num_self_attend_heads: int = 8,
                 cross_attend_widening_factor: int = 1, self_attend_widening_factor: int = 1,
                 dropout_prob: float = 0.0, latent_pos_enc_init_scale: float = 0.02,
                 cross_attention_shape_for_attn: str = "kv", use_query_residual: bool = True):
        super().__init__()
        if num_latent_channels % num_self_attend_heads != 0:
            raise ValueError(f"num_z_channels ({num_latent_channels}) must be divisible by"
                             f" num_self_attend_heads ({num_self_attend_heads}).")
        if num_latent_channels % num_cross_attend_heads != 0:
            raise ValueError(f"num_z_channels ({num_latent_channels}) must be divisible by"
                             f" num_cross_attend_heads ({num_cross_attend_heads}).")
        self._num_blocks = num_blocks
        self.latent_pos_enc = TrainablePositionEncoding(index_dim=num_latents, num_channels=num_latent_channels,
                                                        init_scale=latent_pos_enc_init_scale)
        self.cross_attend = CrossAttention(q_in_channels=num_latent_channels, kv_in_channels=num_input_channels,
                                           dropout_prob=dropout_prob, num_heads=num_cross_attend_heads,
                                           widening_factor=cross_attend_widening_factor,
                                           shape_for_attn=cross_attention_shape_for_attn, qk_channels=qk_channels,
                                           v_channels=v_channels, use_query_residual=use_query_residual)
        self.self_attends = nn.ModuleList()
        for _ in range(num_self_attends_per_block):
            self.self_attends.append(SelfAttention(in_channels=num_latent_channels, num_heads=num_self_attend_heads,
                                                   dropout_prob=dropout_prob, qk_channels=qk_channels,
                                                   v_channels=v_channels,
                                                   widening_factor=self_attend_widening_factor))
        self.key_shard = None  # set by parallel.shard_encoder_keys(): inputs hold this rank's slice of the key axis
        # Encode-once: the reference's chunked decoders call the whole PerceiverIO — encoder
        # included — once per output chunk on the SAME inputs (multimodal_perceiver.py:146-161: 128 calls).  With
        # cache_latents = True the encoder returns the latents of the previous call when the 128-bit content hash of
        # (inputs, latents, mask) equals the previous call's and its parameters are unchanged (opt-in: one read of the
        # input array by pio_hash_words + a 16-byte read-back per call; nothing of the input is kept alive).
        self.cache_latents = False
        self._latent_cache = None
        # Tower LayerNorms folded into the projections: None = automatic (latent arrays of at least
        # engine.FUSE_LN_MIN_ROWS rows, or of engine.FUSE_LN_MIN_CHANNELS channels with fp16 operands, and only while the residual stream passes the numerics check below), True / False
        # force it.  The fused form rounds x to bf16 BEFORE the mean is subtracted, which scales that product's rounding
        # error by sqrt(1 + mean^2 / var) per row: harmless for residual streams centred near zero, not for rows with a
        # large common offset.  So the first fused forward after a parameter change measures max |mean| / std over every
        # row of every residual-stream state (from the statistics the GEMMs produced anyway; one host sync, skipped
        # during CUDA-graph capture) and falls back to the LayerNorm kernels when it exceeds engine.FUSE_LN_MAX_OFFSET.
        self.fuse_layernorm = None
        # arithmetic mode of this module's forward: None = the global engine.PRECISION, or "bf16" / "fp16" / "bf16x3"
        self.precision = None
        self.fused_layernorm_offset = None      # the measured max |mean| / std (None: not measured yet)
        self._fuse_ln_checked = None

    def latents(self, inputs):
        return self.latent_pos_enc(batch_size=inputs.shape[0])

    def forward(self, inputs, latents, *, input_mask=None):
        """inputs fp32 [B, Nk, C_in] (this rank's key slice when `key_shard` is set), latents [B, Nlat, C]."""
        with engine.precision_scope(self.precision):
            return self._forward(inputs, latents, input_mask=input_mask)

    def _forward(self, inputs, latents, *, input_mask=None):
        ops._need_cuda(inputs, latents)
        if isinstance(inputs, torch.Tensor) and inputs.dtype != torch.float32:
            inputs = inputs.float()     # e.g. fp16 activations under the reference's autocast (flow_perceiver.py:129)
        if latents.dtype != torch.float32:
            latents = latents.float()
        if latents.shape[0] == 0 or latents.shape[1] == 0:   # empty batch / no latents: nothing to launch
            return latents.new_empty(latents.shape)
        if isinstance(inputs, PositionedInput) and not engine.fast():
            inputs = inputs.dense()   # the validation precision works on the dense array
        key_mask = None
        row_keep = None
        if input_mask is not None:
            key_mask = input_mask.to(torch.bool)
            if self.key_shard is None:
                any_key = key_mask.any(dim=1, keepdim=True)
                row_keep = any_key.expand(latents.shape[0], latents.shape[1])
            # key-sharded: whether a sample has a valid key on ANY rank comes out of the merged partial sums
            # (pio_combine_args.row_alive), so the exchange stays ONE step
        use_cache = (self.cache_latents and self.key_shard is None and inputs.is_cuda
                     and not torch.cuda.is_current_stream_capturing())
        if use_cache:
            # the reference's wrappers rebuild the preprocessed input array for every chunk, so identity of the tensor
            # object cannot be used: key on the content (a PositionedInput is hashed as its two parts)
            parts = (inputs.features, inputs.pos) if isinstance(inputs, PositionedInput) else (inputs, None)
            lat = latents[0] if latents.dim() == 3 and latents.stride(0) == 0 else latents   # stride-0 batch broadcast
            key = (tuple((p.data_ptr(), p._version) for p in self.parameters()), engine.PRECISION, tuple(latents.shape),
                   ops.content_hash(parts[0], parts[1], lat, input_mask))
            c = self._latent_cache
            if c is not None and c[0] == key:
                return c[1]
        z = self._encode(inputs, latents, key_mask, row_keep)
        if use_cache:
            self._latent_cache = (key, z)
        return z

    def _encode(self, inputs, latents, key_mask, row_keep):
        B, N, C = latents.shape
        layers = [sa for _ in range(self._num_blocks) for sa in self.self_attends]
        fused = None
        want = self.fuse_layernorm
        if want is None:
            want = engine.FUSE_LN and (B * N >= engine.FUSE_LN_MIN_ROWS
                                       or (C >= engine.FUSE_LN_MIN_CHANNELS and engine.PRECISION == "fp16"))
        pver = None
        if want and self.fuse_layernorm is None:
            pver = tuple((p.data_ptr(), p._version) for p in self.parameters())
            if self._fuse_ln_checked is not None and self._fuse_ln_checked[0] == pver and not self._fuse_ln_checked[1]:
                want = False        # these parameters failed the offset check before
        if (want and engine.fast() and layers
                and latents.is_cuda and not any(sa.training and any(p > 0 for p in sa._dropout_probs) for sa in layers)):
            fused = [engine.prepared(sa, "fused", lambda sa=sa: engine.PreparedFusedLayer(sa)) for sa in self.self_attends]
            if not all(pf.usable() for pf in fused):
                fused = None
        if fused is None:
            z, _ = self.cross_attend._forward_factored(latents, inputs, key_mask=key_mask, row_keep=row_keep,
                                                       shard=self.key_shard)
            for self_attend in layers:
                z = self_attend(z)
            return z
        # Latent tower with the LayerNorms folded into the projections: every residual-stream
        # state travels as (fp32 rows, their raw bf16 copy, per-row sum / sum of squares); no LayerNorm kernel runs.
        M = B * N
        stats = torch.empty((2 * len(layers) + 1, M, ops.stats_parts(M, C), 2), dtype=torch.float32, device=latents.device)
        z, zb = self.cross_attend._forward_factored(latents, inputs, key_mask=key_mask, row_keep=row_keep,
                                                    shard=self.key_shard, want_bf16_out=True, stats_out=stats[0])
        x = z.view(M, C)
        # large towers (every producer GEMM on the CTA-pair kernel) carry the stream between their GEMMs as a 16-bit
        # (hi, lo) pair; the tower's input and its output — the latents the caller gets — stay fp32
        split = engine.SPLIT_STREAM and ops.gemm_uses_pair_kernel(M, C)
        x_lo = None
        for i in range(len(layers)):
            pf = fused[i % len(self.self_attends)]
            last = i == len(layers) - 1
            x, zb, x_lo = engine.self_attention_block_fused(pf, x, zb, stats[2 * i], B=B, N=N, st_mid=stats[2 * i + 1],
                                                            st_out=None if last else stats[2 * i + 2], x_lo=x_lo,
                                                            split=split, split_out=split and not last)
        if (pver is not None and (self._fuse_ln_checked is None or self._fuse_ln_checked[0] != pver)
                and not torch.cuda.is_current_stream_capturing()):
            s = stats[:2 * len(layers)].sum(2)                       # [states, M, 2]: (sum, sum of squares) per row
            mean = s[..., 0] / C
            var = (s[..., 1] / C - mean * mean).clamp_min(0.0)
            offset = float((mean.abs() / (var + 1e-5).sqrt()).max())
            self.fused_layernorm_offset = offset
            ok = offset <= engine.FUSE_LN_MAX_OFFSET
            self._fuse_ln_checked = (pver, ok)
            if not ok:
                import warnings
                warnings.warn(f"perceiverio_pytorch_b200: residual-stream rows reach |mean| / std = {offset:.2f} > "
                              f"{engine.FUSE_LN_MAX_OFFSET}; the latent tower keeps its LayerNorm kernels for this model")
                return self._encode(inputs, latents, key_mask, row_keep)
        return x.view(B, N, C)


class _ComposedFinal(nn.Module):
    """final_layer followed by a postprocessor's own Linear (postprocessors.py:176-187 `ClassificationPostprocessor.linear`,
    :200-208 `ProjectionPostprocessor.projection`) as ONE affine map: post(final(x)) = x (Wp Wf)^T + (Wp bf + bp), composed
    in fp64.  Quacks like the nn.Linear the decoder's tail code expects (`weight`, `bias`);
    holds the source layers (and the MLP whose second layer the narrow-head fold absorbs) as sub-modules so that
    engine.prepared() sees every parameter the derived weights depend on.  Lives in the decoder's __dict__, not in its
    module tree: the decoder's state_dict stays the reference's."""

    def __init__(self, final_layer: nn.Linear, post: nn.Linear, mlp: nn.Module):
        super().__init__()
        self.final_layer, self.post, self.mlp = final_layer, post, mlp
        self._ver = None
        self._w = self._b = None

    def _refresh(self):
        ver = tuple((p.data_ptr(), p._version, p.device) for p in (*self.final_layer.parameters(), *self.post.parameters()))
        if ver != self._ver:
            d = torch.float64
            wf, wp = self.final_layer.weight.detach().to(d), self.post.weight.detach().to(d)
            b = wp @ self.final_layer.bias.detach().to(d) if self.final_layer.bias is not None else wp.new_zeros(wp.shape[0])
            if self.post.bias is not None:
                b = b + self.post.bias.detach().to(d)
            self._w, self._b, self._ver = (wp @ wf).float().contiguous(), b.float().contiguous(), ver

    @property
    def weight(self) -> torch.Tensor:
        self._refresh()
        return self._w

    @property
    def bias(self) -> torch.Tensor:
        self._refresh()
        return self._b


class PerceiverDecoder(nn.Module):
    """Cross-attention-based decoder (reference: perceiver.py:110-180)."""

    def __init__(self, query_channels: int, final_project_out_channels: int, num_latent_channels: int = 1024,
                 qk_channels: int = None, v_channels: int = None, use_query_residual: bool = False,
                 output_w_init: str = "lecun_normal", num_heads: int = 1, final_project: bool = True):
        super().__init__()
        self._output_num_channels = final_project_out_channels
        self._output_w_init = output_w_init
        self._use_query_residual = use_query_residual
        self._qk_channels = qk_channels
        self._v_channels = v_channels
        self._final_project = final_project
        self._num_heads = num_heads
        self.query_channels = query_channels
        self.decoding_cross_attn = CrossAttention(q_in_channels=query_channels, kv_in_channels=num_latent_channels,
                                                  dropout_prob=0.0, num_heads=self._num_heads, widening_factor=1,
                                                  shape_for_attn="kv", qk_channels=self._qk_channels,
                                                  v_channels=self._v_channels,
                                                  use_query_residual=self._use_query_residual)
        if self._final_project:
            self.final_layer = nn.Linear(query_channels, self._output_num_channels)
            if self._output_w_init == "lecun_normal":
                lecun_normal_(self.final_layer.weight)
            elif self._output_w_init == "zeros":
                nn.init.constant_(self.final_layer.weight, 0)
            else:
                raise ValueError(f"{self._output_w_init} not supported as output_w_init")
            nn.init.constant_(self.final_layer.bias, 0)
        # arithmetic mode of this module's forward: None = the global engine.PRECISION, or "bf16" / "fp16" / "bf16x3"
        self.precision = None

    def forward(self, query, latents, *, query_mask=None, post_linear: nn.Linear = None):
        """Reference signature (perceiver.py:166) plus `post_linear`: a postprocessor's Linear to apply to the decoder's
        output inside its last projection (N3: `final_layer` and the Linear composed into one map, no intermediate
        [B, Nq, out] array and one GEMM less).  Ignored — the caller applies it — unless the decoder projects."""
        with engine.precision_scope(self.precision):
            return self._forward(query, latents, query_mask=query_mask, post_linear=post_linear)

    def fuses_post_linear(self, post_linear) -> bool:
        return bool(self._final_project and isinstance(post_linear, nn.Linear)
                    and post_linear.in_features == self._output_num_channels)

    def _final(self, post_linear):
        if post_linear is None:
            return self.final_layer, self._output_num_channels
        held = self.__dict__.get("_pio_composed")
        if held is None or held.post is not post_linear or held.final_layer is not self.final_layer:
            held = _ComposedFinal(self.final_layer, post_linear, self.decoding_cross_attn.mlp)
            self.__dict__["_pio_composed"] = held
        return held, post_linear.out_features

    def _forward(self, query, latents, *, query_mask=None, post_linear=None):
        ops._need_cuda(query, latents)
        if query.dtype != torch.float32:
            query = query.float()       # e.g. fp16 activations under the reference's autocast (flow_perceiver.py:129)
        if latents.dtype != torch.float32:
            latents = latents.float()
        row_keep = query_mask.to(torch.bool) if query_mask is not None else None
        if post_linear is not None and not self.fuses_post_linear(post_linear):
            raise ValueError("PerceiverDecoder: post_linear must be an nn.Linear over this decoder's projected outputs")
        fin, n_out = self._final(post_linear) if self._final_project else (None, self._output_num_channels)
        if query.shape[0] == 0 or query.shape[1] == 0:   # empty batch / no output queries: nothing to launch
            return query.new_empty(query.shape[0], query.shape[1], n_out if self._final_project else query.shape[2])
        # the wide final projection consumes bf16 rows: let the MLP's last GEMM write them next to the fp32 result
        want16 = (self._final_project and n_out > 16 and engine.fast()
                  and self.query_channels % 16 == 0)
        tail = None
        if self._final_project and n_out <= 16 and engine.fast():
            # a handful of output channels (optical flow: 322 -> 2): the head is folded into the MLP's second layer and
            # evaluated in fp32 on CUDA cores together with it.  The head is 0.01 % of the FLOPs but its 16-bit rounding
            # alone would cost 1.2e-2 of the 1e-2 error budget in bf16.
            tail = (engine.prepared(self, "tail", lambda: engine.PreparedTail(self.decoding_cross_attn.mlp, fin))
                    if fin is self.final_layer else
                    engine.prepared(fin, "tail", lambda: engine.PreparedTail(self.decoding_cross_attn.mlp, fin)))
        y32, y16 = self.decoding_cross_attn._forward_factored(query, latents, key_mask=None, row_keep=row_keep,
                                                              want_bf16_out=want16, tail=tail)
        if tail is not None:
            return y32
        if not self._final_project:
            return y32 if y32.is_contiguous() else y32.contiguous()   # odd widths carry a 16-byte row pitch inside
        B, Nq, C = y32.shape
        if engine.PRECISION == "bf16x3":
            return validate.final_layer(fin, y32)
        if n_out <= 16:
            # a handful of output channels (optical flow: 322 -> 2): fp32 on CUDA cores.  The head is 0.01 % of the
            # FLOPs but its bf16 rounding alone would cost 1.2e-2 of the 1e-2 error budget.
            bias = fin.bias.detach() if fin.bias is not None else None
            out = ops.linear_f32(y32.view(B * Nq, C), fin.weight.detach(), bias)
            return out.view(B, Nq, -1)
        w = engine.prepared(fin, "w", lambda: (engine._bf16_weight(fin.weight.detach()),
                                               fin.bias.detach().float().contiguous()))
        if y16 is None:
            y16 = ops.layernorm_bf16(y32.view(B * Nq, C), None, None, normalize=False)
        out, _ = ops.linear(y16, C, w[0], n_out, w[1], want_f32=True, want_bf16=False)
        return out.contiguous().view(B, Nq, -1)
