"""Oracle: UpCUNet / CUNet forward from a state_dict (TEST INFRASTRUCTURE).

Functional fp32 restatement of waifu2x/models/cunet.py and the SE block of
nunif/modules/attention.py:29-44.  ``dtype`` lets tests mimic the reference's
CUDA autocast (fp16 convs) when run on a GPU.
"""
import torch
import torch.nn.functional as F


def _lrelu(x):
    return F.leaky_relu(x, 0.1)


def _se(sd, p, x):
    """attention.py:38-44."""
    z = F.adaptive_avg_pool2d(x, 1)
    z = F.relu(F.conv2d(z, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"]))
    z = torch.sigmoid(F.conv2d(z, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"]))
    return x * z


def _unet_conv(sd, p, x, se):
    """cunet.py:10-28."""
    z = _lrelu(F.conv2d(x, sd[p + ".conv.0.weight"], sd[p + ".conv.0.bias"]))
    z = _lrelu(F.conv2d(z, sd[p + ".conv.2.weight"], sd[p + ".conv.2.bias"]))
    if se:
        z = _se(sd, p + ".seblock", z)
    return z


def unet1(sd, p, x, deconv):
    """cunet.py:55-67."""
    x1 = _unet_conv(sd, p + ".conv1", x, False)
    x2 = _lrelu(F.conv2d(x1, sd[p + ".conv1_down.weight"], sd[p + ".conv1_down.bias"], stride=2))
    x2 = _unet_conv(sd, p + ".conv2", x2, True)
    x2 = _lrelu(F.conv_transpose2d(x2, sd[p + ".conv2_up.weight"], sd[p + ".conv2_up.bias"], stride=2))
    x1 = x1[:, :, 4:-4, 4:-4]
    x3 = _lrelu(F.conv2d(x1 + x2, sd[p + ".conv3.weight"], sd[p + ".conv3.bias"]))
    if deconv:
        return F.conv_transpose2d(x3, sd[p + ".conv_bottom.weight"], sd[p + ".conv_bottom.bias"],
                                  stride=2, padding=3)
    return F.conv2d(x3, sd[p + ".conv_bottom.weight"], sd[p + ".conv_bottom.bias"])


def unet2(sd, p, x):
    """cunet.py:99-121."""
    x1 = _unet_conv(sd, p + ".conv1", x, False)
    x2 = _lrelu(F.conv2d(x1, sd[p + ".conv1_down.weight"], sd[p + ".conv1_down.bias"], stride=2))
    x2 = _unet_conv(sd, p + ".conv2", x2, True)
    x3 = _lrelu(F.conv2d(x2, sd[p + ".conv2_down.weight"], sd[p + ".conv2_down.bias"], stride=2))
    x3 = _unet_conv(sd, p + ".conv3", x3, True)
    x3 = _lrelu(F.conv_transpose2d(x3, sd[p + ".conv3_up.weight"], sd[p + ".conv3_up.bias"], stride=2))
    x2 = x2[:, :, 4:-4, 4:-4]
    x4 = _unet_conv(sd, p + ".conv4", x2 + x3, True)
    x4 = _lrelu(F.conv_transpose2d(x4, sd[p + ".conv4_up.weight"], sd[p + ".conv4_up.bias"], stride=2))
    x1 = x1[:, :, 16:-16, 16:-16]
    x5 = _lrelu(F.conv2d(x1 + x4, sd[p + ".conv5.weight"], sd[p + ".conv5.bias"]))
    return F.conv2d(x5, sd[p + ".conv_bottom.weight"], sd[p + ".conv_bottom.bias"])


def cunet_forward(sd, x, upscale, no_clip=False):
    """UpCUNet.forward / CUNet.forward eval path (cunet.py:149-163, 183-197)."""
    z1 = unet1(sd, "unet1", x, deconv=upscale)
    if not no_clip:
        z1 = torch.clamp(z1, 0., 1.)
    z2 = unet2(sd, "unet2", z1)
    z1 = z1[:, :, 20:-20, 20:-20]
    return torch.clamp(z1 + z2, 0., 1.)


UPCUNET = dict(scale=2, offset=36, blend_size=0)   # cunet.py:144
CUNET = dict(scale=1, offset=28, blend_size=0)     # cunet.py:178
