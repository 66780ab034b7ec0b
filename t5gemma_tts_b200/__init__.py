"""t5gemma_tts_b200 -- B200-native (sm_100a) engine for the T5Gemma-TTS token-generation hot path.

Host side of the drop-in boundary: `T5GemmaVoiceEngine` mirrors the
reference's `T5GemmaVoiceModel` / `T5GemmaVoiceForConditionalGeneration` surface that
`inference_tts_utils.inference_one_sample` touches (`.config`, `.args`, `.eval()`, `.to()`,
`.inference_tts(...)`), and drives hand-written CUDA through the C ABI in include/t5gtts.h.
There is no CPU fallback: importing the engine without libt5gtts.so raises.
"""
from .config import EngineConfig
from .engine import T5GemmaVoiceEngine, GenerationRequest
from .lib import load_library, T5GError

__all__ = ["EngineConfig", "T5GemmaVoiceEngine", "GenerationRequest", "load_library", "T5GError"]
