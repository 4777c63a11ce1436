"""Fused per-frame stereo pipelines (what iw3/utils.py:292-391 apply_divergence +
:430-487 postprocess_image do for the SBS / anaglyph outputs of the BASELINE configs)."""
from .backward_warp import apply_divergence_grid_sample
from .forward_warp import apply_divergence_forward_warp
from .depth_scaler import minmax_normalize
from .dilation import dilate_edge, edge_dilation_is_enabled
from ._common import COMPOSE_SBS, COMPOSE_ANAGLYPH


def stereo_sbs(c, depth, divergence=2.0, convergence=0.5, method="forward_fill", mapper="none",
               edge_dilation=0, synthetic_view="both", anaglyph=None):
    """c: B,3,H,W frames; depth: B,1,h,w raw model output (larger = nearer).
    dilate_edge -> per-frame min/max -> mapper -> warp -> SBS (B,3,H,2W) or anaglyph (B,3,H,W)."""
    if edge_dilation_is_enabled(edge_dilation):
        depth = dilate_edge(depth, edge_dilation)
    depth = minmax_normalize(depth, mapper=mapper)
    if method in {"forward", "forward_fill"}:
        if anaglyph is not None:
            from .anaglyph import apply_anaglyph_redcyan
            l, r = apply_divergence_forward_warp(c, depth, divergence, convergence, method=method,
                                                 synthetic_view=synthetic_view, width_base=False)
            return apply_anaglyph_redcyan(l, r, anaglyph)
        return apply_divergence_forward_warp(c, depth, divergence, convergence, method=method,
                                             synthetic_view=synthetic_view, width_base=False, compose=COMPOSE_SBS)
    if method in {"grid_sample", "backward"}:
        compose = COMPOSE_SBS if anaglyph is None else COMPOSE_ANAGLYPH
        if anaglyph not in (None, "dubois"):
            raise NotImplementedError("fused anaglyph epilogue supports dubois only")
        return apply_divergence_grid_sample(c, depth, divergence, convergence, synthetic_view, compose=compose)
    raise ValueError(f"method {method} is not on the B200 hot path")
