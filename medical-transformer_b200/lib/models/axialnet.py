"""B200-native drop-in for the reference's lib/models/axialnet.py.

Same public names, constructor signatures, parameter/buffer names, shapes, registration
order and initialisation (so `torch.manual_seed(s); gated(img_size=128)` yields the
reference's weights bit for bit, and checkpoints load in both directions), but every
AxialAttention* forward is ONE call into the hand-written sm_100a kernels behind the C ABI
(medt_b200.axial_attention -> libmedt_b200.so); convolutions stay on cuDNN; the LoGo local
branch runs all 16 patches as one on-device batch with per-patch BN statistics.

Reference map (file:line in /root/reference/lib/models/axialnet.py):
  AxialAttention :19-97 | AxialAttention_dynamic :99-193 | AxialAttention_wopos :195-258
  AxialBlock :262-302 | AxialBlock_dynamic :304-344 | AxialBlock_wopos :346-391
  ResAxialAttentionUNet :397-507 | medt_net :509-711 | factories :714-728
There is no CPU path: forward on a non-CUDA tensor raises (BASELINE north_star: "no CPU
fallback")."""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from medt_b200.axial_op import axial_attention
from medt_b200.shell_ops import bn_act, conv1x1 as conv1x1_op, conv2d, up2x_relu_add

from .utils import qkv_transform

PATCH = 32          # the reference hard-codes 32-px local patches on a 4x4 grid (:661-664,:700)
GRID = 4


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


# ----------------------------------------------------------------------------- attention ops
class _AxialAttentionBase(nn.Module):
    """Shared constructor of the three attention variants.  `_variant` selects which terms
    exist: "gated" (rel-pos + gates), "pos" (rel-pos), "nopos" (qk only)."""
    _variant = None

    def __init__(self, in_planes, out_planes, groups=8, kernel_size=56, stride=1, bias=False, width=False):
        assert (in_planes % groups == 0) and (out_planes % groups == 0)
        super().__init__()
        if in_planes != out_planes or groups != 8:
            raise NotImplementedError("the sm_100a kernels cover in_planes == out_planes, groups == 8 "
                                      "(the only configuration the reference ever builds)")
        self.in_planes, self.out_planes, self.groups = in_planes, out_planes, groups
        self.group_planes = out_planes // groups
        self.kernel_size, self.stride, self.bias, self.width = kernel_size, stride, bias, width
        self._stat_groups = 1        # >1 only inside the batched LoGo local branch
        has_pos = self._variant != "nopos"

        self.qkv_transform = qkv_transform(in_planes, out_planes * 2, kernel_size=1, stride=1, padding=0, bias=False)
        self.bn_qkv = nn.BatchNorm1d(out_planes * 2)
        self.bn_similarity = nn.BatchNorm2d(groups * 3 if has_pos else groups)
        self.bn_output = nn.BatchNorm1d(out_planes * 2 if has_pos else out_planes)
        if self._variant == "gated":                      # frozen until train.py:169-171 flips them on
            self.f_qr = nn.Parameter(torch.tensor(0.1), requires_grad=False)
            self.f_kr = nn.Parameter(torch.tensor(0.1), requires_grad=False)
            self.f_sve = nn.Parameter(torch.tensor(0.1), requires_grad=False)
            self.f_sv = nn.Parameter(torch.tensor(1.0), requires_grad=False)
        if has_pos:
            self.relative = nn.Parameter(torch.randn(self.group_planes * 2, kernel_size * 2 - 1), requires_grad=True)
            span = torch.arange(kernel_size)
            # redundant for the kernels (they index the Toeplitz table directly) but part of the checkpoint
            self.register_buffer('flatten_index', (span.unsqueeze(1) - span.unsqueeze(0) + kernel_size - 1).view(-1))
        if stride > 1:
            self.pooling = nn.AvgPool2d(stride, stride=stride)      # folded into the kernel's epilogue
        self.reset_parameters()

    def forward(self, x):
        return axial_attention(x, self, self._stat_groups)

    def reset_parameters(self):
        self.qkv_transform.weight.data.normal_(0, math.sqrt(1. / self.in_planes))
        if self._variant != "nopos":
            nn.init.normal_(self.relative, 0., math.sqrt(1. / self.group_planes))


class AxialAttention(_AxialAttentionBase):
    _variant = "pos"


class AxialAttention_dynamic(_AxialAttentionBase):
    _variant = "gated"


class AxialAttention_wopos(_AxialAttentionBase):
    _variant = "nopos"


# ----------------------------------------------------------------------------- residual blocks
class _AxialBlockBase(nn.Module):
    expansion = 2
    _attn = None
    _extra_conv1 = False     # AxialBlock_wopos carries an unused 1x1 conv (reference :358)

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1,
                 base_width=64, dilation=1, norm_layer=None, kernel_size=56):
        super().__init__()
        if norm_layer is None:
            norm_layer = nn.BatchNorm2d
        width = int(planes * (base_width / 64.))
        self.conv_down = conv1x1(inplanes, width)
        if self._extra_conv1:
            self.conv1 = nn.Conv2d(width, width, kernel_size=1)
        self.bn1 = norm_layer(width)
        self.hight_block = self._attn(width, width, groups=groups, kernel_size=kernel_size)
        self.width_block = self._attn(width, width, groups=groups, kernel_size=kernel_size, stride=stride, width=True)
        self.conv_up = conv1x1(width, planes * self.expansion)
        self.bn2 = norm_layer(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self._stat_groups = 1

    def forward(self, x):
        g = self._stat_groups
        idt, side, cur = x, None, None
        if self.downsample is not None:
            # the projection shortcut (1x1 conv + BN) is independent of the main branch until the residual add:
            # fork it onto a second stream so it (and, through autograd, its backward) overlaps the attention chain
            side = _fork_stream(x) if x.is_cuda else None
            if side is not None:
                cur = torch.cuda.current_stream(x.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    idt = bn_act(conv1x1_op(x, self.downsample[0]), self.downsample[1], g, relu=False)
            else:
                idt = bn_act(conv1x1_op(x, self.downsample[0]), self.downsample[1], g, relu=False)
        out = bn_act(conv1x1_op(x, self.conv_down), self.bn1, g, relu=True)
        out = self.width_block(self.hight_block(out))
        if os.environ.get("MEDT_NO_SHELL_FUSION"):
            out = conv1x1_op(F.relu(out), self.conv_up)
        else:
            out = conv1x1_op(out, self.conv_up, relu_in=True)
        if side is not None:
            cur.wait_stream(side)
            idt.record_stream(cur)
        return bn_act(out, self.bn2, g, relu=True, residual=idt)


_FORK_STREAMS = {}


def _fork_stream(x):
    """One forked stream per (device, current stream): the shortcut branches of blocks running on different
    streams (MedT's global / local branches) must not share one, or they would serialise each other."""
    if os.environ.get("MEDT_NO_BRANCH_STREAM"):
        return None
    cur = torch.cuda.current_stream(x.device)
    key = (x.device.index, cur.cuda_stream)
    st = _FORK_STREAMS.get(key)
    if st is None:
        st = _FORK_STREAMS[key] = torch.cuda.Stream(device=x.device)
    return st


class AxialBlock(_AxialBlockBase):
    _attn = AxialAttention


class AxialBlock_dynamic(_AxialBlockBase):
    _attn = AxialAttention_dynamic


class AxialBlock_wopos(_AxialBlockBase):
    _attn = AxialAttention_wopos
    _extra_conv1 = True


# ----------------------------------------------------------------------------- shells
def _up_relu(conv, x, skip=None):
    """relu(bilinear x2 (conv(x))) [+ skip] -- one decoder step (reference :493-501), one fused kernel."""
    if conv.bias is None or os.environ.get("MEDT_NO_SHELL_FUSION"):
        return up2x_relu_add(conv2d(x, conv), skip)
    # the conv runs without its bias; the fused kernel adds it (the bilinear taps sum to 1) and reduces its gradient
    y = conv2d(x, conv, with_bias=False)
    return up2x_relu_add(y, skip, conv.bias)


class _ShellBase(nn.Module):
    def _check_dilation_arg(self, r):
        if r is None:
            r = [False, False, False]
        if len(r) != 3:
            raise ValueError("replace_stride_with_dilation should be None "
                             "or a 3-element tuple, got {}".format(r))
        return r

    def _make_layer(self, block, planes, blocks, kernel_size=56, stride=1, dilate=False):
        norm_layer = self._norm_layer
        previous_dilation = self.dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(conv1x1(self.inplanes, planes * block.expansion, stride),
                                       norm_layer(planes * block.expansion))
        seq = [block(self.inplanes, planes, stride, downsample, groups=self.groups, base_width=self.base_width,
                     dilation=previous_dilation, norm_layer=norm_layer, kernel_size=kernel_size)]
        self.inplanes = planes * block.expansion
        if stride != 1:
            kernel_size = kernel_size // 2
        seq += [block(self.inplanes, planes, groups=self.groups, base_width=self.base_width, dilation=self.dilation,
                      norm_layer=norm_layer, kernel_size=kernel_size) for _ in range(1, blocks)]
        return nn.Sequential(*seq)

    @staticmethod
    def _stem(x, convs, bns, groups=1):
        for conv, bn in zip(convs, bns):
            x = bn_act(conv2d(x, conv), bn, groups, relu=True)
        return x


class ResAxialAttentionUNet(_ShellBase):
    """Axial-attention U-Net: conv stem -> 4 axial layers -> 5 conv/upsample decoder steps."""

    def __init__(self, block, layers, num_classes=2, zero_init_residual=True, groups=8, width_per_group=64,
                 replace_stride_with_dilation=None, norm_layer=None, s=0.125, img_size=128, imgchan=3):
        super().__init__()
        self._norm_layer = norm_layer = nn.BatchNorm2d if norm_layer is None else norm_layer
        self.inplanes = int(64 * s)
        self.dilation = 1
        rsd = self._check_dilation_arg(replace_stride_with_dilation)
        self.groups, self.base_width = groups, width_per_group
        self.conv1 = nn.Conv2d(imgchan, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.conv2 = nn.Conv2d(self.inplanes, 128, kernel_size=3, stride=1, padding=1, bias=False)
        self.conv3 = nn.Conv2d(128, self.inplanes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1, self.bn2, self.bn3 = norm_layer(self.inplanes), norm_layer(128), norm_layer(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(block, int(128 * s), layers[0], kernel_size=(img_size // 2))
        self.layer2 = self._make_layer(block, int(256 * s), layers[1], stride=2, kernel_size=(img_size // 2), dilate=rsd[0])
        self.layer3 = self._make_layer(block, int(512 * s), layers[2], stride=2, kernel_size=(img_size // 4), dilate=rsd[1])
        self.layer4 = self._make_layer(block, int(1024 * s), layers[3], stride=2, kernel_size=(img_size // 8), dilate=rsd[2])
        self.decoder1 = nn.Conv2d(int(1024 * 2 * s), int(1024 * 2 * s), kernel_size=3, stride=2, padding=1)
        self.decoder2 = nn.Conv2d(int(1024 * 2 * s), int(1024 * s), kernel_size=3, stride=1, padding=1)
        self.decoder3 = nn.Conv2d(int(1024 * s), int(512 * s), kernel_size=3, stride=1, padding=1)
        self.decoder4 = nn.Conv2d(int(512 * s), int(256 * s), kernel_size=3, stride=1, padding=1)
        self.decoder5 = nn.Conv2d(int(256 * s), int(128 * s), kernel_size=3, stride=1, padding=1)
        self.adjust = nn.Conv2d(int(128 * s), num_classes, kernel_size=1, stride=1, padding=0)
        self.soft = nn.Softmax(dim=1)        # constructed, never applied (reference :440,:502)

    def _forward_impl(self, x, features=False):
        x = self._stem(x, (self.conv1, self.conv2, self.conv3), (self.bn1, self.bn2, self.bn3))
        x1 = self.layer1(x)
        x2 = self.layer2(x1)
        x3 = self.layer3(x2)
        x4 = self.layer4(x3)
        x = _up_relu(self.decoder1, x4, x4)
        x = _up_relu(self.decoder2, x, x3)
        x = _up_relu(self.decoder3, x, x2)
        x = _up_relu(self.decoder4, x, x1)
        x = _up_relu(self.decoder5, x)
        return x if features else self.adjust(F.relu(x))

    def forward(self, x):
        return self._forward_impl(x)

    def forward_features(self, x):
        """everything up to (not including) `self.adjust(F.relu(.))` -- medt_b200.TrainStep fuses that tail with the loss"""
        return self._forward_impl(x, features=True)


class medt_net(_ShellBase):
    """MedT / LoGo: a shallow global branch on the whole image plus a deep local branch on a
    4x4 grid of 32-px patches, merged before the head."""

    def __init__(self, block, block_2, layers, num_classes=2, zero_init_residual=True, groups=8, width_per_group=64,
                 replace_stride_with_dilation=None, norm_layer=None, s=0.125, img_size=128, imgchan=3, patch_size=None):
        super().__init__()
        # patch_size=None keeps the reference's hard-coded 32-px patches (bug-compatible: at img_size != 128
        # only the top-left 128x128 gets local features and `logo` cannot run); patch_size=img_size//4 is the
        # generalisation SURVEY 8(f)-1 asks for (identical to the reference at 128 px).
        self._patch = PATCH if patch_size is None else int(patch_size)
        self._norm_layer = norm_layer = nn.BatchNorm2d if norm_layer is None else norm_layer
        self.inplanes = int(64 * s)
        self.dilation = 1
        rsd = self._check_dilation_arg(replace_stride_with_dilation)
        self.groups, self.base_width = groups, width_per_group
        # ---- global branch
        self.conv1 = nn.Conv2d(imgchan, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.conv2 = nn.Conv2d(self.inplanes, 128, kernel_size=3, stride=1, padding=1, bias=False)
        self.conv3 = nn.Conv2d(128, self.inplanes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1, self.bn2, self.bn3 = norm_layer(self.inplanes), norm_layer(128), norm_layer(self.inplanes)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = self._make_layer(block, int(128 * s), layers[0], kernel_size=(img_size // 2))
        self.layer2 = self._make_layer(block, int(256 * s), layers[1], stride=2, kernel_size=(img_size // 2), dilate=rsd[0])
        self.decoder4 = nn.Conv2d(int(512 * s), int(256 * s), kernel_size=3, stride=1, padding=1)
        self.decoder5 = nn.Conv2d(int(256 * s), int(128 * s), kernel_size=3, stride=1, padding=1)
        self.adjust = nn.Conv2d(int(128 * s), num_classes, kernel_size=1, stride=1, padding=0)
        self.soft = nn.Softmax(dim=1)
        # ---- local branch.  NB: the reference builds its stem with the *current* self.inplanes,
        # i.e. the global layer2's 64 output channels (:557-566), not int(64*s).
        self.conv1_p = nn.Conv2d(imgchan, self.inplanes, kernel_size=7, stride=2, padding=3, bias=False)
        self.conv2_p = nn.Conv2d(self.inplanes, 128, kernel_size=3, stride=1, padding=1, bias=False)
        self.conv3_p = nn.Conv2d(128, self.inplanes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1_p, self.bn2_p, self.bn3_p = norm_layer(self.inplanes), norm_layer(128), norm_layer(self.inplanes)
        self.relu_p = nn.ReLU(inplace=True)
        img_size_p = img_size // 4
        self.layer1_p = self._make_layer(block_2, int(128 * s), layers[0], kernel_size=(img_size_p // 2))
        self.layer2_p = self._make_layer(block_2, int(256 * s), layers[1], stride=2, kernel_size=(img_size_p // 2), dilate=rsd[0])
        self.layer3_p = self._make_layer(block_2, int(512 * s), layers[2], stride=2, kernel_size=(img_size_p // 4), dilate=rsd[1])
        self.layer4_p = self._make_layer(block_2, int(1024 * s), layers[3], stride=2, kernel_size=(img_size_p // 8), dilate=rsd[2])
        self.decoder1_p = nn.Conv2d(int(1024 * 2 * s), int(1024 * 2 * s), kernel_size=3, stride=2, padding=1)
        self.decoder2_p = nn.Conv2d(int(1024 * 2 * s), int(1024 * s), kernel_size=3, stride=1, padding=1)
        self.decoder3_p = nn.Conv2d(int(1024 * s), int(512 * s), kernel_size=3, stride=1, padding=1)
        self.decoder4_p = nn.Conv2d(int(512 * s), int(256 * s), kernel_size=3, stride=1, padding=1)
        self.decoder5_p = nn.Conv2d(int(256 * s), int(128 * s), kernel_size=3, stride=1, padding=1)
        self.decoderf = nn.Conv2d(int(128 * s), int(128 * s), kernel_size=3, stride=1, padding=1)
        self.adjust_p = nn.Conv2d(int(128 * s), num_classes, kernel_size=1, stride=1, padding=0)   # unused (:588)
        self.soft_p = nn.Softmax(dim=1)
        # every BN of the local branch sees one patch per call in the reference -> 16 statistic groups
        self._local_groups = GRID * GRID
        for layer in (self.layer1_p, self.layer2_p, self.layer3_p, self.layer4_p):
            for mod in layer.modules():
                if hasattr(mod, "_stat_groups"):
                    mod._stat_groups = self._local_groups

    def _local_branch(self, xin):
        n = xin.shape[0]
        PATCH = self._patch
        side = PATCH * GRID
        # (N,3,128,128) -> (16N,3,32,32), patch-major in the reference's (i, j) loop order
        xp = xin[:, :, :side, :side].reshape(n, -1, GRID, PATCH, GRID, PATCH).permute(2, 4, 0, 1, 3, 5)
        xp = xp.reshape(GRID * GRID * n, -1, PATCH, PATCH)
        g = self._local_groups
        xp = self._stem(xp, (self.conv1_p, self.conv2_p, self.conv3_p), (self.bn1_p, self.bn2_p, self.bn3_p), g)
        x1 = self.layer1_p(xp)
        x2 = self.layer2_p(x1)
        x3 = self.layer3_p(x2)
        x4 = self.layer4_p(x3)
        xp = _up_relu(self.decoder1_p, x4, x4)
        xp = _up_relu(self.decoder2_p, xp, x3)
        xp = _up_relu(self.decoder3_p, xp, x2)
        xp = _up_relu(self.decoder4_p, xp, x1)
        xp = _up_relu(self.decoder5_p, xp)
        c = xp.shape[1]
        return xp.reshape(GRID, GRID, n, c, PATCH, PATCH).permute(2, 3, 0, 4, 1, 5).reshape(n, c, side, side)

    def _forward_impl(self, x, features=False):
        xin = x
        # The global and the local branch are independent until the merge: the local branch runs on a
        # second CUDA stream (autograd replays its backward on that stream too), so the two chains of
        # microsecond-scale kernels overlap on the 148 SMs.  Fork/join is capturable in the step graph.
        cur = torch.cuda.current_stream(x.device) if x.is_cuda else None
        side = self._side_stream(x.device) if x.is_cuda else None
        if side is not None:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                loc = self._local_branch(xin)
        x = self._stem(x, (self.conv1, self.conv2, self.conv3), (self.bn1, self.bn2, self.bn3))
        x1 = self.layer1(x)
        x2 = self.layer2(x1)
        x = _up_relu(self.decoder4, x2, x1)
        x = _up_relu(self.decoder5, x)
        if side is not None:
            cur.wait_stream(side)
            loc.record_stream(cur)
        else:
            loc = self._local_branch(xin)       # raises: no CPU path
        side_len = self._patch * GRID
        x_loc = x.clone()                       # outside the 128x128 patch grid x_loc stays == x (:658,:700)
        x_loc[:, :, :side_len, :side_len] = loc
        x = F.relu(conv2d(x + x_loc, self.decoderf))
        return x if features else self.adjust(F.relu(x))

    def _side_stream(self, device):
        # module-level cache keyed by (device, current stream): nn.DataParallel rebuilds its replicas on every forward, so
        # a per-instance cache would create (and destroy) a stream per replica and iteration
        if os.environ.get("MEDT_NO_BRANCH_STREAM"):
            return None
        dev = torch.device(device)
        key = ("local", dev.index, torch.cuda.current_stream(dev).cuda_stream)
        st = _FORK_STREAMS.get(key)
        if st is None:
            st = _FORK_STREAMS[key] = torch.cuda.Stream(device=dev)
        return st

    def forward(self, x):
        return self._forward_impl(x)

    def forward_features(self, x):
        """everything up to (not including) `self.adjust(F.relu(.))` -- medt_b200.TrainStep fuses that tail with the loss"""
        return self._forward_impl(x, features=True)


# ----------------------------------------------------------------------------- factories
def axialunet(pretrained=False, **kwargs):
    return ResAxialAttentionUNet(AxialBlock, [1, 2, 4, 1], s=0.125, **kwargs)


def gated(pretrained=False, **kwargs):
    return ResAxialAttentionUNet(AxialBlock_dynamic, [1, 2, 4, 1], s=0.125, **kwargs)


def MedT(pretrained=False, **kwargs):
    return medt_net(AxialBlock_dynamic, AxialBlock_wopos, [1, 2, 4, 1], s=0.125, **kwargs)


def logo(pretrained=False, **kwargs):
    return medt_net(AxialBlock, AxialBlock, [1, 2, 4, 1], s=0.125, **kwargs)
