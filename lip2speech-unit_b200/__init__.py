"""B200-native multi_input_vocoder generator forward (lip2speech-unit hot path).

The directory name carries a hyphen (it follows the reference repository's
name), so it is loaded through ``__graft_entry__.load_package()`` /
``importlib`` under the module name ``lip2speech_unit_b200``; alternatively put
this directory on ``sys.path`` and ``from models_multi_input import
MelCodeGenerator`` exactly like the reference's inference scripts do.
"""
from . import _cabi  # noqa: F401
from .models_multi_input import AttrDict, CodeGenerator, MelCodeGenerator  # noqa: F401
from .dispatch import shard_utterances, chunk_plan, vocode_long, HostPipeline, MultiGpuVocoder  # noqa: F401
from . import hand_off  # noqa: F401

__all__ = ["MelCodeGenerator", "CodeGenerator", "AttrDict", "shard_utterances", "chunk_plan", "vocode_long", "HostPipeline", "MultiGpuVocoder", "hand_off"]
