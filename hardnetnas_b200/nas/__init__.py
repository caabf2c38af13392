"""NAS-derived descriptor nets (hardnetNAS/fbnet_building_blocks + supernet stem/head) on the B200 path."""
from .fbnet_builder import PRIMITIVES, ChannelShuffle, ConvBNRelu, Flatten, IRFBlock, Identity, SEModule  # noqa: F401
from .fbnet_modeldef import MODEL_ARCH  # noqa: F401
from .lookup_table import CANDIDATE_BLOCKS, SEARCH_SPACE2, LookUpTable  # noqa: F401
from .descriptor_net import SampledDescriptorNet  # noqa: F401
