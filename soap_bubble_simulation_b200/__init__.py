"""soap_bubble_simulation_b200 -- B200-native drop-in for the thin-film drainage + interference path of
pjt222/soap_bubble_simulation.  Host-side mirrors of the reference API over the C ABI in include/sbx.h."""
from . import _lib, shader_code
from ._lib import (F32, F64, OUT_RGB_F32, OUT_RGBA8, PREC_F32, PREC_F64, PREC_F64_STRICT, RGB_EXACT3,
                   RGB_EXACT3_LINEAR, RGB_LUT, RGB_NONE, RGB_SCHLICK_TWO_BEAM, RGB_SCHLICK_WALL, RGB_SPECTRAL, SEM_CPU, SEM_WGSL, SbxError)
from .config import (BubbleParameters, ConfigError, EnvironmentParameters, FluidParameters,
                     SimulationConfig)
from .drainage import DrainageSimulator, ThicknessField, View, camera_position
from .gpu_drainage import DrainageParams, GPUDrainageSimulator
from .headless import (BubbleUniform, Camera, ExportError, HeadlessRenderPipeline, export_frame,
                       render_drainage_frame)
from .interference import (LUT_ANGLE_SAMPLES, LUT_MAX_THICKNESS_NM, LUT_THICKNESS_SAMPLES,
                           FresnelCoefficients, InterferenceCalculator, RgbColor,
                           generate_interference_lut, generate_spectral_lut, thin_film_interference)

__all__ = [n for n in dir() if not n.startswith("_")]
