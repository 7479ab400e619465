"""nova_pointcloud_b200 -- B200-native (sm_100a) diffusion-head sampling for NOVA point clouds.

Drop-in for ONE hot path of zailaiyiwan123/NOVA_pointcloud: the ``DiffusionMLP`` head driven by
the flow-matching Euler scheduler over every point token of every autoregressive set, plus a
Chamfer-distance scorer.  Python here is the reference-facing surface only; the arithmetic lives
in ``lib/libnova_b200.so`` (hand-written CUDA, C ABI in ``include/nova_b200.h``) and is reached
through ``torch.ops.nova_b200``.  CUDA only -- nothing falls back to the CPU.
"""

from ._lib import LIB_PATH, NovaError  # noqa: F401
from . import ops  # noqa: F401  (registers torch.ops.nova_b200.*)
from .modules import AdaLayerNormZero, DiffusionBlock, DiffusionMLP, PatchEmbed, Projector, TimeCondEmbed  # noqa: F401
from .schedulers import FlowMatchEulerDiscreteScheduler, FlowMatchEulerDiscreteSchedulerOutput  # noqa: F401
from .registry import IMAGE_DECODERS, POINT_CLOUD_DECODERS, Registry  # noqa: F401
from .pipeline import (  # noqa: F401
    GuidanceScaler,
    NOVAPointCloudGenerationPipeline,
    NOVAPointCloudPipelineOutput,
    NOVATrainPointCloudPipeline,
    HostSampler,
    denoise,
    gather_shards,
    generate_sets,
    sample_sharded,
    shard_range,
    standard_point_cloud_generation,
)
from .chamfer import (  # noqa: F401
    chamfer_distance,
    chamfer_nn,
    compute_chamfer_distance,
    compute_emd,
    earth_mover_distance,
    emd_approx,
    emd_matching,
    robust_emd,
    dist_chamfer,
    robust_chamfer_distance,
)
from .geometry import (  # noqa: F401
    compute_local_density,
    density_target_size,
    dynamic_partition,
    farthest_point_sampling,
    feature_aware_interpolation,
    knn,
)
from .training import get_losses  # noqa: F401
from . import partition, synth  # noqa: F401

__version__ = "0.1.0"
