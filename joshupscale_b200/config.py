"""Model configuration for the per-frame recurrent upscaling graph.

Mirrors the keyword arguments of the reference's model factory functions:
  - get_flow_autoencoder  (scripts/training/models.py:334-339, defaults 364-365)
  - get_flow_resnet       (scripts/training/models.py:257-263)
  - get_generator_resnet  (scripts/training/models.py:484-491)
  - get_inference_model   (scripts/training/models.py:680-689, padding 735-741)

The reference does not pin the hyper-parameters of its shipped engines
(model_psp.trt / model_psp_fast.trt / model_ps2.trt / model_ps2_fast.trt are
file names only, obs_plugin/src/filter.cc:138-143); the presets below are the
function defaults plus the frame sizes from SURVEY.md section 8.
"""

from __future__ import annotations

import dataclasses
from typing import List, Tuple

SCALE = 4  # kScale, core/src/tensorrt_backend.cc:27
BN_EPS = 1e-3  # keras.layers.BatchNormalization default epsilon
LRELU_SLOPE = 0.3  # keras 3 LeakyReLU default negative_slope


@dataclasses.dataclass(frozen=True)
class ModelConfig:
    frame_height: int = 270
    frame_width: int = 480
    flow_pad_factor: int = 8
    flow_arch: str = "autoencoder"  # "autoencoder" | "resnet"
    flow_num_inputs: int = 4
    flow_filters: Tuple[int, ...] = (32, 64, 128, 256, 128, 64, 32)
    flow_resnet_filters: int = 64
    flow_resnet_blocks: int = 10
    flow_activation: str = "relu"  # "relu" | "lrelu"
    gen_filters: int = 64
    gen_blocks: int = 24
    gen_activation: str = "relu"
    normalize_brightness: bool = False

    @property
    def padded_height(self) -> int:
        f = self.flow_pad_factor
        return (self.frame_height + f - 1) // f * f if f else self.frame_height

    @property
    def padded_width(self) -> int:
        f = self.flow_pad_factor
        return (self.frame_width + f - 1) // f * f if f else self.frame_width

    @property
    def pad_top(self) -> int:
        # ZeroPadding2D((pad//2, pad - pad//2)), scripts/training/models.py:780-789
        return (self.padded_height - self.frame_height) // 2

    @property
    def pad_left(self) -> int:
        return (self.padded_width - self.frame_width) // 2

    @property
    def out_height(self) -> int:
        return self.frame_height * SCALE

    @property
    def out_width(self) -> int:
        return self.frame_width * SCALE

    def validate(self) -> None:
        if self.flow_arch not in ("autoencoder", "resnet"):
            raise ValueError(f"unknown flow_arch {self.flow_arch}")
        if self.flow_arch == "autoencoder":
            n = len(self.flow_filters) // 2
            div = 1 << n
            if self.padded_height % div or self.padded_width % div:
                raise ValueError(
                    "padded frame must be divisible by 2**(len(filters)//2)")
        if self.frame_height < 2 or self.frame_width < 2:
            raise ValueError("frame too small")

    def flow_gmacs(self) -> float:
        """Algorithmic multiply-accumulates of the flow net (true Cin)."""
        ph, pw = self.padded_height, self.padded_width
        cin = 3 * self.flow_num_inputs
        macs = 0
        if self.flow_arch == "autoencoder":
            f = self.flow_filters
            n = len(f) // 2
            h, w = ph, pw
            for i in range(n):
                macs += h * w * 9 * (cin * f[i] + f[i] * f[i])
                cin = f[i]
                h //= 2
                w //= 2
            for i in range(n, 2 * n):
                macs += h * w * 9 * (cin * f[i] + f[i] * f[i])
                cin = f[i]
                h *= 2
                w *= 2
            if len(f) % 2:
                macs += h * w * 9 * cin * f[-1]
                cin = f[-1]
            macs += h * w * 9 * cin * 32
        else:
            nf = self.flow_resnet_filters
            macs += ph * pw * 9 * cin * nf
            macs += ph * pw * 9 * nf * nf * 2 * self.flow_resnet_blocks
            macs += ph * pw * nf * 32
        return macs / 1e9

    def gen_gmacs(self) -> float:
        h, w, nf = self.frame_height, self.frame_width, self.gen_filters
        macs = h * w * 9 * 51 * nf
        macs += h * w * 9 * nf * nf * 2 * self.gen_blocks
        macs += h * w * nf * 32 * 4
        macs += (2 * h) * (2 * w) * 32 * 3 * 4
        return macs / 1e9

    def gflop_per_frame(self) -> float:
        return 2.0 * (self.flow_gmacs() + self.gen_gmacs())


@dataclasses.dataclass(frozen=True)
class OutputFilter:
    """Output temporal filter + scene-cut gate that the reference bakes into the
    exported graph; fields and defaults are the CLI arguments of
    scripts/inference/onnx/frame_moving_avg.py:53-87.  Stored in the model
    container as the float32[8] tensor `meta/frame_moving_avg`
    (weights.with_output_filter) and executed by csrc/kernels/frame_filter.cu."""
    strength: float = 0.25
    window: int = 0  # scene detection window in output pixels, 0 = global
    threshold: float = 0.1
    gain: float = 0.0  # 0 = sign function, otherwise tanh(gain * ...)
    norm: str = "l1"  # "l1" | "l2"
    limit: bool = False  # clip the warped previous output to [-0.5, 0.5] first
    luma_normalize: bool = False

    def as_vector(self):
        if self.norm.lower() not in ("l1", "l2"):
            raise ValueError(f"unknown norm {self.norm}")
        if self.window < 0:
            raise ValueError("window must be >= 0")
        return [1.0, float(self.strength), float(self.window), float(self.threshold), float(self.gain),
                1.0 if self.norm.lower() == "l2" else 0.0, 1.0 if self.limit else 0.0,
                1.0 if self.luma_normalize else 0.0]


def preset(name: str) -> ModelConfig:
    """Named configurations."""
    presets = {
        # function defaults at the PSP frame size
        "psp_quality": ModelConfig(),
        # builder-defined (the reference ships only the file name)
        "psp_fast": ModelConfig(gen_blocks=8),
        # PS2 size inferred from obs_plugin/data/mask.png (1920x1440)
        "ps2_quality": ModelConfig(frame_height=360, frame_width=480),
        "ps2_fast": ModelConfig(frame_height=360, frame_width=480, gen_blocks=8),
        # the alternative flow architecture (get_flow_resnet defaults, models.py:257-263) at PSP size
        "psp_resnet": ModelConfig(flow_arch="resnet", flow_pad_factor=0),
        # small shapes for CPU-side parity tests (exercise asymmetric padding)
        "tiny": ModelConfig(frame_height=21, frame_width=27, gen_blocks=2,
                            flow_filters=(8, 16, 16, 32, 16, 16, 8)),
        "small": ModelConfig(frame_height=46, frame_width=72, gen_blocks=3),
        "small_bright": ModelConfig(frame_height=46, frame_width=72, gen_blocks=2,
                                    normalize_brightness=True, flow_num_inputs=3),
        "small_resnet": ModelConfig(frame_height=46, frame_width=72, gen_blocks=2,
                                    flow_arch="resnet", flow_resnet_blocks=2,
                                    flow_pad_factor=0),
    }
    return presets[name]


def layer_names(cfg: ModelConfig) -> List[str]:
    """Names of all parametrised layers, in execution order."""
    names: List[str] = []
    if cfg.flow_arch == "autoencoder":
        n = len(cfg.flow_filters) // 2
        for i in range(2 * n):
            for j in (1, 2):
                names.append(f"flow/block_{i + 1}/conv_{j}")
                names.append(f"flow/block_{i + 1}/bn_{j}")
        if len(cfg.flow_filters) % 2:
            names += ["flow/conv_1", "flow/bn_1"]
        names.append("flow/conv_2")
    else:
        names += ["flow/conv_1", "flow/bn_1"]
        for i in range(cfg.flow_resnet_blocks):
            for j in (1, 2):
                names.append(f"flow/block_{i + 1}/conv_{j}")
                names.append(f"flow/block_{i + 1}/bn_{j}")
        names.append("flow/conv_2")
    names += ["generator/conv_1", "generator/bn_1"]
    for i in range(cfg.gen_blocks):
        for j in (1, 2):
            names.append(f"generator/block_{i + 1}/conv_{j}")
            names.append(f"generator/block_{i + 1}/bn_{j}")
    names += ["generator/conv_trans_1", "generator/bn_2", "generator/conv_trans_2"]
    return names
