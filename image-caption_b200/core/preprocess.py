"""Drop-in for the region feature extractor of the reference's core/preprocess.py:26-62 (`ResnetExtractor`): the
ResNet-101 trunk (torchvision `resnet101` children[:9] = conv1, bn1, relu, maxpool, layer1..4, avgpool) over the
224 x 224 crops of an image's regions -> one 2048-d feature per region (the stage before the caption
model).  Same constructor-free interface (`forward(x) -> ndarray [N, 2048]`, `transform(image)`, `transforms`,
`image_size`); the parameters carry torchvision's names (`submodule.state_dict()` == the reference's
`nn.Sequential(*children[:9]).state_dict()`), so pretrained weights load unchanged when they are available -- offline the
trunk is initialised like torchvision's (`weights=None`).

Execution (libicap, include/icap.h): activations are NHWC matrices [N*H*W, C]; every convolution is ONE icap_gemm
(tcgen05 / TMA in bf16 mode, true-fp32 SIMT in fp32 mode) over the activation matrix itself (1x1) or over the patch
matrix gathered by icap_im2col_nhwc (7x7 stem, 3x3, strided 1x1); BatchNorm + ReLU + residual add are icap_bn_scale_shift
+ icap_bn_act.  The reference never switches its extractor to eval() (preprocess.py:35-40), so BatchNorm uses the
statistics of the batch of crops; `.eval()` gives the running-statistics form.  The detector in front of it (YOLOv5,
data/detect_for_preprocess.py) needs pretrained weights and stays out of scope: `image_feature_YOLOv5` takes the boxes.
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

from core.config import DEVICE, ENCODE_DIM_FEATURES, ENCODE_DIM_POSITIONS, NUM_OBJECT

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)
import icap_loader  # noqa: E402

_pkg = icap_loader.load()
_N = _pkg._native
call, F32, BF16 = _N.call, _N.F32, _N.BF16

LAYERS = ((64, 3, 1), (128, 4, 2), (256, 23, 2), (512, 3, 2))      # (width, blocks, stride of the first block): resnet101
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class _Seq(nn.Module):
    """Numeric-named container: reproduces the dotted names of nn.Sequential / torchvision's Bottleneck."""


def _conv_param(cout, cin, k):
    w = torch.empty(cout, cin, k, k)
    nn.init.kaiming_normal_(w, mode="fan_out", nonlinearity="relu")          # torchvision ResNet.__init__
    return nn.Parameter(w)


def _add_bn(mod, name, c):
    bn = _Seq()
    bn.register_parameter("weight", nn.Parameter(torch.ones(c)))
    bn.register_parameter("bias", nn.Parameter(torch.zeros(c)))
    bn.register_buffer("running_mean", torch.zeros(c))
    bn.register_buffer("running_var", torch.ones(c))
    bn.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
    mod.add_module(name, bn)


def _add_conv(mod, name, cout, cin, k):
    conv = _Seq()
    conv.register_parameter("weight", _conv_param(cout, cin, k))
    mod.add_module(name, conv)


def _build_trunk():
    trunk = _Seq()
    _add_conv(trunk, "0", 64, 3, 7)
    _add_bn(trunk, "1", 64)
    cin = 64
    for li, (width, blocks, _stride) in enumerate(LAYERS):
        layer = _Seq()
        for b in range(blocks):
            blk = _Seq()
            _add_conv(blk, "conv1", width, cin, 1)
            _add_bn(blk, "bn1", width)
            _add_conv(blk, "conv2", width, width, 3)
            _add_bn(blk, "bn2", width)
            _add_conv(blk, "conv3", 4 * width, width, 1)
            _add_bn(blk, "bn3", 4 * width)
            if b == 0:
                ds = _Seq()
                _add_conv(ds, "0", 4 * width, cin, 1)
                _add_bn(ds, "1", 4 * width)
                blk.add_module("downsample", ds)
            layer.add_module(str(b), blk)
            cin = 4 * width
        trunk.add_module(str(4 + li), layer)
    return trunk


class ResnetExtractor(nn.Module):
    def __init__(self):
        super(ResnetExtractor, self).__init__()
        self.norm_mean = [0.485, 0.456, 0.406]
        self.norm_std = [0.229, 0.224, 0.225]
        self.size = 224
        self.precision = os.environ.get("ICAP_PRECISION", "bf16")      # "bf16" (tcgen05) | "fp32" (parity mode)
        self.submodule = _build_trunk()
        if torch.cuda.is_available():
            self.submodule.to(DEVICE)
        self._packed = {}
        self.launches = 0

    # ------------------------------------------------------------------ reference interface
    def forward(self, x):
        """x [N, 3, 224, 224] normalised crops -> ndarray [N, 2048] (preprocess.py:42-46)."""
        with torch.no_grad():
            return self.features(x).cpu().numpy()

    def transform(self, image):
        """BGR uint8 image (cv2) -> [1, 3, 224, 224] float tensor: bicubic resize, RGB, /255, normalise
        (preprocess.py:48-55; ToTensor + Normalize restated in numpy)."""
        import cv2
        resized = cv2.resize(image, (self.size, self.size), interpolation=cv2.INTER_CUBIC)
        rgb = cv2.cvtColor(resized, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0
        rgb = (rgb - np.asarray(self.norm_mean, np.float32)) / np.asarray(self.norm_std, np.float32)
        return torch.from_numpy(np.ascontiguousarray(rgb.transpose(2, 0, 1))).unsqueeze(0)

    @property
    def transforms(self):
        return self.transform

    @property
    def image_size(self):
        return self.size

    # ------------------------------------------------------------------ execution on libicap
    def _s(self):
        return torch.cuda.current_stream(self._dev).cuda_stream

    def _weight(self, mod, k):
        """conv weight [Cout, Cin, k, k] -> GEMM operand [Cout, ld] in (ky, kx, c) order, ld = k*k*Cin rounded up to 8."""
        w = mod.weight
        key = id(w)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == w._version and hit[1].dtype == self._tdt:
            return hit[1]
        cout, cin = w.shape[0], w.shape[1]
        K = k * k * cin
        ld = (K + 7) // 8 * 8
        packed = torch.zeros(cout, ld, dtype=self._tdt, device=self._dev)
        packed[:, :K] = w.detach().permute(0, 2, 3, 1).reshape(cout, K).to(self._tdt)
        self._packed[key] = (w._version, packed)
        return packed

    def _conv(self, x, shape, mod, k, stride, pad):
        n, h, w_, c = shape
        wmat = self._weight(mod, k)
        cout, ld = wmat.shape
        ho, wo = (h + 2 * pad - k) // stride + 1, (w_ + 2 * pad - k) // stride + 1
        m = n * ho * wo
        if k == 1 and stride == 1:
            a = x
        else:
            a = torch.empty(m, ld, dtype=self._tdt, device=self._dev)
            call("icap_im2col_nhwc", self._act, x.data_ptr(), n, h, w_, c, k, k, stride, pad, a.data_ptr(), ld, self._s())
            self.launches += 1
        y = torch.empty(m, cout, dtype=self._tdt, device=self._dev)
        call("icap_gemm", self._act, 1, 1, m, cout, ld, a.data_ptr(), ld, wmat.data_ptr(), ld, y.data_ptr(), cout, self._act,
             None, _N.EPI_B_STATIC if self._act == BF16 else 0, None, 0, 0, 1, self._s())
        self.launches += 1
        return y, (n, ho, wo, cout)

    def _bn(self, x, bn, relu, residual=None):
        m, c = x.shape
        train = int(self.training)
        call("icap_bn_scale_shift", self._act, x.data_ptr(), m, c, self._sums.data_ptr(), bn.weight.data_ptr(),
             bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), BN_MOMENTUM, BN_EPS, train,
             self._scale.data_ptr(), self._shift.data_ptr(), self._s())
        y = torch.empty_like(x)
        call("icap_bn_act", self._act, x.data_ptr(), m, c, self._scale.data_ptr(), self._shift.data_ptr(),
             residual.data_ptr() if residual is not None else None, int(relu), y.data_ptr(), self._s())
        self.launches += 3 if train else 2
        return y

    @torch.no_grad()
    def features(self, x):
        """[N, 3, H, W] (CPU or CUDA) -> CUDA tensor [N, 2048] fp32."""
        t = self.submodule
        first = getattr(t, "0").weight
        if not first.is_cuda:
            raise _N.IcapError("ResnetExtractor runs on sm_100a GPUs only (no CPU fallback): move it to the GPU first")
        self._dev = first.device
        self._act = BF16 if self.precision == "bf16" else F32
        self._tdt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        self._sums = torch.zeros(2 * 2048, dtype=torch.float64, device=self._dev)
        self._scale = torch.empty(2048, dtype=torch.float32, device=self._dev)
        self._shift = torch.empty(2048, dtype=torch.float32, device=self._dev)
        n, c, h, w = x.shape
        assert c == 3
        # every weight operand is packed BEFORE the first launch: the GEMMs may fetch their weight tiles before their grid
        # dependency resolves (ICAP_EPI_B_STATIC), which is only safe when no packing copy runs right before them
        bns = []
        for name, mod in t.named_modules():
            if hasattr(mod, "weight") and mod.weight.dim() == 4:
                self._weight(mod, mod.weight.shape[-1])
            elif hasattr(mod, "num_batches_tracked"):
                bns.append(mod.num_batches_tracked)
        # NCHW fp32 -> NHWC in the compute dtype (layout glue of the 3-channel input only)
        a = x.to(self._dev, torch.float32).permute(0, 2, 3, 1).contiguous().to(self._tdt)
        y, shp = self._conv(a.view(n * h * w, 3), (n, h, w, 3), getattr(t, "0"), 7, 2, 3)
        y = self._bn(y, getattr(t, "1"), relu=True)
        n_, h_, w_, c_ = shp
        ho, wo = (h_ + 2 - 3) // 2 + 1, (w_ + 2 - 3) // 2 + 1
        p = torch.empty(n_ * ho * wo, c_, dtype=self._tdt, device=self._dev)
        call("icap_maxpool_nhwc", self._act, y.data_ptr(), n_, h_, w_, c_, 3, 2, 1, p.data_ptr(), self._s())
        self.launches += 1
        y, shp = p, (n_, ho, wo, c_)
        for li, (_width, blocks, stride) in enumerate(LAYERS):
            layer = getattr(t, str(4 + li))
            for b in range(blocks):
                blk = getattr(layer, str(b))
                s = stride if b == 0 else 1
                identity = y
                o, shp1 = self._conv(y, shp, blk.conv1, 1, 1, 0)
                o = self._bn(o, blk.bn1, relu=True)
                o, shp2 = self._conv(o, shp1, blk.conv2, 3, s, 1)          # torchvision: the stride sits on the 3x3
                o = self._bn(o, blk.bn2, relu=True)
                o, shp3 = self._conv(o, shp2, blk.conv3, 1, 1, 0)
                if hasattr(blk, "downsample"):
                    identity, _ = self._conv(y, shp, getattr(blk.downsample, "0"), 1, s, 0)
                    identity = self._bn(identity, getattr(blk.downsample, "1"), relu=False)
                y = self._bn(o, blk.bn3, relu=True, residual=identity)
                shp = shp3
        n_, h_, w_, c_ = shp
        out = torch.empty(n_, c_, dtype=torch.float32, device=self._dev)
        call("icap_avgpool_nhwc", self._act, y.data_ptr(), n_, h_ * w_, c_, out.data_ptr(), self._s())
        self.launches += 1
        if self.training:
            torch._foreach_add_(bns, 1)                     # nn.BatchNorm2d bookkeeping
        return out


def image_feature_YOLOv5(image, boxes_xyxy, positions, num_obj=NUM_OBJECT, extractor=None):
    """Feature / position arrays of ONE image in the reference's layout (preprocess.py:91-138): region 0 is the whole
    image with position [0, 0, 1, 1, 0 x 80], then one row per detected box, zero-padded to num_obj + 1 rows.  The boxes
    (pixel xyxy) and their 84-d position rows come from the detector, which is out of scope here (YOLOv5 needs
    pretrained weights): pass them in."""
    extractor = extractor or ResnetExtractor()
    crops = [extractor.transform(image)]
    for x0, y0, x1, y1 in boxes_xyxy[:num_obj]:
        crops.append(extractor.transform(image[int(y0):int(y1), int(x0):int(x1)]))
    feats = extractor(torch.cat(crops))
    pos = [[0, 0, 1, 1] + [0] * (ENCODE_DIM_POSITIONS - 4)] + [list(p) for p in positions[:num_obj]]
    pos += [[0] * ENCODE_DIM_POSITIONS] * (num_obj + 1 - len(pos))
    if feats.shape[0] < num_obj + 1:
        feats = np.concatenate([feats, np.zeros((num_obj + 1 - feats.shape[0], ENCODE_DIM_FEATURES))])
    return np.asarray(feats).astype(float), np.asarray(pos).astype(float), np.asarray(boxes_xyxy)
