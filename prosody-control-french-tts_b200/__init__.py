"""prosody_b200 — B200-native (sm_100a CUDA) implementation of the prosody-extraction hot path of
hi-paris/Prosody-Control-French-TTS (AudioPipeline.measure_prosody_and_build_ssml, Code/audioPipeline.py:261-711).

    batch      Extractor / Units: batched get_median_pitch / get_lufs / get_part_duration over the C ABI
    legacy     batched calculate_pitch_segment / _calculate_loudness of the DataFrame pipeline
    _native    ctypes binding of libprosody_b200.so (include/prosody_b200.h)
    build      nvcc build of the library (in-tree)
"""
from . import _native
from .batch import Extractor, Units, intensity_plan, part_durations, pitch_frame_times, pitch_params, pitch_plan

__all__ = ["Extractor", "Units", "pitch_params", "pitch_plan", "part_durations", "intensity_plan", "pitch_frame_times", "_native"]
