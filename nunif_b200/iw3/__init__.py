"""B200-native mirror of the iw3 hot-path callables (reference: iw3/*.py).

Same names, argument meaning and error behaviour as the reference functions so
that ``iw3.utils.apply_divergence`` / ``postprocess_image`` can import these in
place of the originals.  Every function requires CUDA tensors
and dispatches to hand-written sm_100a kernels through the C ABI; there is no
CPU path.
"""
from .backward_warp import apply_divergence_grid_sample  # noqa: F401
from .forward_warp import apply_divergence_forward_warp  # noqa: F401
from .dilation import dilate_edge, edge_dilation_parse, edge_dilation_is_enabled  # noqa: F401
from .depth_scaler import minmax_normalize, depth_mapper, EMAMinMaxScaler  # noqa: F401
from .base_depth_model import BaseDepthModel  # noqa: F401
from .anaglyph import apply_anaglyph_redcyan  # noqa: F401
from .stereo import stereo_sbs  # noqa: F401
from .frames import hwc_to_chw_float, chw_float_to_hwc  # noqa: F401
from .depth_anything_preprocess import batch_preprocess, preprocess_size  # noqa: F401
from .depth_anything_model import DepthAnythingModel, DepthAnythingNet, batch_infer  # noqa: F401
from . import zoedepth_preprocess  # noqa: F401
from .zoedepth_model import ZoeDepthModel, ZoeDepthNet  # noqa: F401
from .row_flow import RowFlowV3, MLBW, apply_divergence_nn_LR, apply_divergence_nn_delta, apply_divergence_nn_delta_weight  # noqa: F401
from .depth_aa import DepthAA  # noqa: F401
from .postprocess import postprocess_image, postprocess_padding, resize_bicubic_aa, equirectangular_projection  # noqa: F401
from .utils import apply_divergence, process_image  # noqa: F401
