"""Input-side glue of the encoder kept on the device.

The reference's preprocessors build the encoder input as ``cat([features, broadcast(position table)], -1)``
(io_processors/preprocessors.py:180-199): for the ImageNet-pixels recipe that is 3 pixel channels next to 258 Fourier
channels which are the same for every sample — 3.35 GB of fp32 per 64-image batch of which 38 MB are information, and
the table is rebuilt on the CPU and copied to the device on every forward (preprocessors.py:187-188,
position_encoding.py:173-183).  `PositionedInput` carries the two parts separately; `PerceiverEncoder.forward` accepts
it in place of the dense array and normalises it with `pio_layernorm_concat_bf16`, so the concatenated array is never
materialised.  The table is the reference's own (computed once by the reference's position-encoding module, or by
`fourier_position_table`, a restatement of position_encoding.py:19-89) and stays resident on the device.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch


class PositionedInput:
    """Stands for ``torch.cat([features, pos[None].expand(B, -1, -1)], dim=-1)``.

    features: fp32 [B, N, Cf], any strides (e.g. ``img.movedim(-3, -1).reshape(B, H * W, C)`` of an NCHW image is a view);
    pos: fp32 [N, Cp], contiguous, the same for every sample.
    Only what the reference's glue asks of the encoder input is provided: ``shape`` / ``device`` / ``dtype`` (batch size
    for the latent and query arrays, perceiver.py:302-305) and slicing along the index axis (``restructure``,
    perceiver.py:370-387)."""

    def __init__(self, features: torch.Tensor, pos: torch.Tensor):
        if features.dim() != 3 or pos.dim() != 2 or pos.shape[0] != features.shape[1]:
            raise ValueError(f"PositionedInput: features [B, N, Cf] and pos [N, Cp] expected, got "
                             f"{tuple(features.shape)} and {tuple(pos.shape)}")
        if features.device != pos.device:
            raise ValueError("PositionedInput: features and pos must be on the same device")
        self.features = features.float() if features.dtype != torch.float32 else features
        self.pos = pos.float().contiguous()

    @property
    def shape(self):
        b, n, cf = self.features.shape
        return torch.Size((b, n, cf + self.pos.shape[1]))

    @property
    def device(self):
        return self.features.device

    @property
    def dtype(self):
        return torch.float32

    @property
    def is_cuda(self):
        return self.features.is_cuda

    def dim(self):
        return 3

    def __getitem__(self, idx):
        """Slicing along the batch and index axes (``x[:, a:b]``), as `restructure` does."""
        if not isinstance(idx, tuple):
            idx = (idx,)
        if len(idx) > 2 or not all(isinstance(i, slice) for i in idx):
            raise TypeError("PositionedInput supports slicing along the batch and index axes only")
        b = idx[0]
        n = idx[1] if len(idx) > 1 else slice(None)
        if n.step not in (None, 1):
            raise TypeError("PositionedInput: strided index slices are not supported")
        return PositionedInput(self.features[b, n], self.pos[n])

    def dense(self) -> torch.Tensor:
        """The concatenated fp32 array the reference would have built."""
        b = self.features.shape[0]
        return torch.cat([self.features, self.pos[None].expand(b, -1, -1)], dim=-1)


def fourier_position_table(index_dims: Sequence[int], num_bands: int, max_resolution: Optional[Sequence[int]] = None,
                           concat_pos: bool = True, sine_only: bool = False, device=None) -> torch.Tensor:
    """[prod(index_dims), C_pos] Fourier features of a linear position grid in [-1, 1]^d — what
    ``FourierPositionEncoding(index_dims, num_bands, ...)(batch_size=None)`` returns
    (position_encoding.py:19-67 `generate_fourier_features`, :70-89 `build_linear_positions`, :173-183)."""
    max_resolution = tuple(max_resolution or index_dims)
    ranges = [torch.linspace(-1.0, 1.0, steps=n, dtype=torch.float32) for n in index_dims]
    grid = torch.stack(torch.meshgrid(*ranges, indexing="ij"), dim=-1)
    pos = grid.reshape(-1, len(index_dims))
    freq = torch.stack([torch.linspace(1.0, res / 2, steps=num_bands) for res in max_resolution], dim=0)
    per_pos = (pos[:, :, None] * freq[None, :, :]).reshape(pos.shape[0], -1)
    if sine_only:
        feats = torch.sin(math.pi * per_pos)
    else:
        feats = torch.cat([torch.sin(math.pi * per_pos), torch.cos(math.pi * per_pos)], dim=-1)
    if concat_pos:
        feats = torch.cat([pos, feats], dim=-1)
    return feats.to(device) if device is not None else feats


def positioned_image_input(preprocessor, images: torch.Tensor, pos=None) -> Optional[PositionedInput]:
    """`PositionedInput` for a reference ``ImagePreprocessor`` (io_processors/preprocessors.py:57-258) and a batch of
    images already on the device, or None when the preprocessor's configuration is not a plain concatenation of
    per-sample features with a batch-invariant table (``concat_or_add_pos == "add"``, extra position MLPs).  The
    features are produced by the preprocessor's own layers (:216-253); the table by its own position-encoding module,
    once per (module, device) when ``pos`` is None and the encoding has no parameters."""
    if getattr(preprocessor, "_concat_or_add_pos", None) != "concat" or getattr(preprocessor, "_n_extra_pos_mlp", 0) != 0:
        return None
    prep = preprocessor._prep_type
    x = images
    if prep in ("conv", "conv1x1"):
        has_t = x.dim() == 5
        if has_t:
            b, t = x.shape[:2]
            x = x.view(b * t, *x.shape[2:])
        x = preprocessor.convnet(x) if prep == "conv" else preprocessor.convnet_1x1(x)
        x = x.movedim(-3, -1)
        if has_t:
            x = x.view(b, t, *x.shape[1:])
    elif prep == "pixels":
        x = x.movedim(-3, -1)
        sd, td = preprocessor._spatial_downsample, preprocessor._temporal_downsample
        if x.dim() == 4:
            x = x[:, ::sd, ::sd]
        elif x.dim() == 5:
            x = x[:, ::td, ::sd, ::sd]
        else:
            raise ValueError("Unsupported data format for pixels.")
    elif prep == "patches":
        # the preprocessor module's own space_to_depth (processor_utils.py:21-40) and optional projection (:243-249)
        import sys
        s2d = getattr(sys.modules.get(type(preprocessor).__module__), "space_to_depth", None)
        if s2d is None:
            return None
        x = s2d(x.movedim(-3, -1), temporal_block_size=preprocessor._temporal_downsample,
                spatial_block_size=preprocessor._spatial_downsample)
        if x.ndim == 5 and x.shape[1] == 1:
            x = torch.squeeze(x, dim=1)
        if preprocessor._conv_after_patching:
            x = preprocessor._conv_after_patch_layer(x)
    else:
        return None
    batch = x.shape[0]
    n = 1
    for d in preprocessor.index_dims:
        n *= int(d)
    feats = x.reshape(batch, n, -1)
    enc = preprocessor._positional_encoding
    # the encodings are batch-invariant by construction (position_encoding.py:119-121, :173-184 use pos[0] only)
    static = pos is None and not any(True for _ in enc.parameters())
    cache = enc.__dict__.setdefault("_pio_position_table", {}) if static else None   # lives and dies with the module
    table = cache.get(str(feats.device)) if static else None
    if table is None:
        with torch.no_grad():
            table = enc(batch_size=1, pos=pos)[0].to(feats.device).float().contiguous()
        if static:
            cache[str(feats.device)] = table
    return PositionedInput(feats, table)


def _query_reads_inputs(output_query) -> bool:
    """Whether a reference output query hands the preprocessed inputs themselves back as (part of) the decoder query:
    `BasicQuery.forward` returns `inputs` when it has no position encoding (output_queries.py:73-76 — the optical-flow
    recipe's `FlowQuery`, :129-139), and `PerceiverIO.decoder_query` then reshapes / concatenates that tensor
    (perceiver.py:353-360), which needs the dense array."""
    return bool(getattr(output_query, "_concat_preprocessed_input", False)
                and getattr(output_query, "_position_encoding", None) is None)


def _post_linear(posts, decoder):
    """The Linear a single '__default' postprocessor applies first, if the decoder can absorb it (N3)."""
    if not posts or list(posts.keys()) != ["__default"] or not hasattr(decoder, "fuses_post_linear"):
        return None
    post = posts["__default"]
    kind = type(post).__name__
    lin = None
    if kind == "ClassificationPostprocessor" and getattr(post, "_project", False):
        lin = getattr(post, "linear", None)
    elif kind == "ProjectionPostprocessor":
        lin = getattr(post, "projection", None)
    return lin if lin is not None and decoder.fuses_post_linear(lin) else None


def perceiver_io_forward(perceiver, inputs: torch.Tensor, *, subsampled_output_points=None, pos=None, input_mask=None,
                         query_mask=None, only_needed_queries: bool = False, fuse_postprocessor: bool = True):
    """`PerceiverIO.forward` (perceiver.py:287-325) for a single image modality with the input glue fused: the
    preprocessor's features and position table reach the encoder as a `PositionedInput`.  Falls back to the module's
    own forward whenever the configuration is not covered (several modalities, channel padding, modality masking).

    only_needed_queries (off by default): the classification wrapper decodes 1000 output
    queries and its postprocessor keeps query 0 only (postprocessors.py:187).  Decoder rows do not interact (each query
    attends over the latents and goes through the MLP and the final projection on its own), so decoding just the kept
    row gives the same logits for 1/1000 of the decoder work.

    fuse_postprocessor (N3, on by default): `ClassificationPostprocessor.linear` / `ProjectionPostprocessor.projection`
    (postprocessors.py:176-187, :200-208) is composed with the decoder's `final_layer` — post(final(x)) is one affine map
    — so the [B, Nq, out] intermediate and one GEMM disappear; what is left of the postprocessor (keeping query 0) is
    applied here."""
    own = type(perceiver).forward     # the class's forward: `perceiver.forward` may be this very function (install.py)
    mp = perceiver._multi_preprocessor
    preps = getattr(mp, "_preprocessors", None) if mp is not None else None
    if (type(inputs) is not torch.Tensor or pos is not None or preps is None or list(preps.keys()) != ["__default"]
            or mp.padding_embeddings is not None or mp._mask_probs is not None
            or type(preps["__default"]).__name__ != "ImagePreprocessor"):
        return own(perceiver, inputs, subsampled_output_points=subsampled_output_points, pos=pos,
                   input_mask=input_mask, query_mask=query_mask)
    pin = positioned_image_input(preps["__default"], inputs)
    if pin is None:
        return own(perceiver, inputs, subsampled_output_points=subsampled_output_points, pos=pos,
                   input_mask=input_mask, query_mask=query_mask)
    sizes = {"__default": pin.shape[1]}
    encoder_query = perceiver._encoder.latents(pin)
    # queries that pass the preprocessed inputs through (optical flow: the 182,528 inputs ARE the decoder queries) get the
    # dense array — it is the decoder's operand there anyway; the encoder still takes the two parts separately
    query_inputs = pin.dense() if any(_query_reads_inputs(q) for q in perceiver._output_queries.values()) else pin
    decoder_query, query_sizes = perceiver.decoder_query(query_inputs, sizes, {"__default": pin.features},
                                                         subsampled_points=subsampled_output_points)
    latents = perceiver._encoder(pin, encoder_query, input_mask=input_mask)
    posts = perceiver._output_postprocessors
    if (only_needed_queries and posts and list(posts.keys()) == ["__default"]
            and type(posts["__default"]).__name__ == "ClassificationPostprocessor" and decoder_query.shape[1] >= 1):
        decoder_query = decoder_query[:, :1]
        query_mask = None if query_mask is None else query_mask[:, :1]
        query_sizes = {"__default": 1}
    # N3: a classification / projection postprocessor starts with its own Linear on the decoder's projected outputs
    # (postprocessors.py:176-187, :200-208); the decoder composes it with final_layer into one map
    post_linear = _post_linear(posts, perceiver._decoder) if fuse_postprocessor else None
    if post_linear is not None:
        outputs = perceiver._decoder(decoder_query, latents, query_mask=query_mask, post_linear=post_linear)
        kind = type(posts["__default"]).__name__
        return outputs[:, 0, :] if kind == "ClassificationPostprocessor" else outputs
    outputs = perceiver._decoder(decoder_query, latents, query_mask=query_mask)
    if perceiver._output_postprocessors:
        if type(outputs) is torch.Tensor:
            index, split = 0, {}
            for modality in sorted(query_sizes.keys()):
                split[modality] = outputs[:, index:index + query_sizes[modality]]
                index += query_sizes[modality]
            outputs = split
        outputs = {modality: post(outputs[modality], pos=None, modality_sizes=None)
                   for modality, post in perceiver._output_postprocessors.items()}
    if type(outputs) is not torch.Tensor and list(outputs.keys()) == ["__default"]:
        outputs = outputs["__default"]
    return outputs
