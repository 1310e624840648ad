"""GLSL snippets of the interference model -- mirror of ``interference::shader_code``
(src/physics/interference.rs:525-610): four string constants plus ``complete_shader_code()``.

The reference never executes these strings (it only checks that the function names are present,
interference.rs:783-790); they document the f32 shader form of the calculator.  The snippets below are
written for this repository with the same function names, argument order and arithmetic (denominator
guard ``max(den, 1e-4)``, pi literal 3.14159265, gamma 1/2.2), so a maintainer can paste them where the
reference's were used.  ``tests/test_shader_code.py`` compiles them as C++ with a small GLSL prelude and
checks them against the oracle's calculator, so unlike the reference's they are pinned numerically.
"""

# Snell: cos(theta_t) from cos(theta_i); total internal reflection gives 0 (interference.rs:197-211)
SNELLS_LAW_GLSL = """
// cos of the refracted angle inside the film (Snell's law); 0.0 on total internal reflection
float calculateTransmissionAngleCos(float cosIncident, float nMedium, float nFilm) {
    float sinI = sqrt(1.0 - cosIncident * cosIncident);
    float sinT = sinI * (nMedium / nFilm);
    if (abs(sinT) > 1.0) {
        return 0.0;
    }
    return sqrt(1.0 - sinT * sinT);
}
"""

# unpolarised Fresnel reflectance (interference.rs:266-298)
FRESNEL_GLSL = """
// mean of the s- and p-polarised power reflectances at the medium/film interface
float calculateFresnelReflectance(float cosIncident, float nMedium, float nFilm) {
    float cosT = calculateTransmissionAngleCos(cosIncident, nMedium, nFilm);
    float a = nMedium * cosIncident;
    float b = nFilm * cosT;
    float c = nFilm * cosIncident;
    float d = nMedium * cosT;
    float rs = (a - b) / (a + b);
    float rp = (c - d) / (c + d);
    return 0.5 * (rs * rs + rp * rp);
}
"""

# multi-bounce reflected intensity at one wavelength (interference.rs:226-254, 312-353)
INTERFERENCE_INTENSITY_GLSL = """
// reflected intensity of a film of `thickness` (same unit as `wavelength`) seen at cosIncident
float calculateInterferenceIntensity(float thickness, float cosIncident, float wavelength, float nMedium, float nFilm) {
    if (thickness <= 0.0) {
        return 0.0;
    }
    float cosT = calculateTransmissionAngleCos(cosIncident, nMedium, nFilm);
    // optical path difference; the half wavelength is the phase jump at the denser medium
    float opd = 2.0 * nFilm * thickness * cosT + 0.5 * wavelength;
    float phase = 2.0 * 3.14159265 * opd / wavelength;
    float R = calculateFresnelReflectance(cosIncident, nMedium, nFilm);
    float c = cos(phase);
    float num = 2.0 * R * (1.0 - c);
    float den = 1.0 + R * R - 2.0 * R * c;
    return clamp(num / max(den, 0.0001), 0.0, 1.0);
}
"""

# RGB at 650 / 532 / 450 nm, linear and gamma 1/2.2 (interference.rs:380-426)
INTERFERENCE_COLOR_GLSL = """
// linear RGB of the film at the three primaries' wavelengths (nm), in air
vec3 calculateInterferenceColor(float thickness, float cosIncident, float nFilm) {
    const float N_AIR = 1.0;
    const float WL_RED = 650.0;
    const float WL_GREEN = 532.0;
    const float WL_BLUE = 450.0;
    return vec3(calculateInterferenceIntensity(thickness, cosIncident, WL_RED, N_AIR, nFilm),
                calculateInterferenceIntensity(thickness, cosIncident, WL_GREEN, N_AIR, nFilm),
                calculateInterferenceIntensity(thickness, cosIncident, WL_BLUE, N_AIR, nFilm));
}

// display colour: gamma 1/2.2 applied per channel
vec3 calculateInterferenceColorGamma(float thickness, float cosIncident, float nFilm) {
    return pow(calculateInterferenceColor(thickness, cosIncident, nFilm), vec3(1.0 / 2.2));
}
"""


def complete_shader_code() -> str:
    """shader_code::complete_shader_code (interference.rs:601-609): the four snippets in dependency order."""
    return "\n".join([SNELLS_LAW_GLSL, FRESNEL_GLSL, INTERFERENCE_INTENSITY_GLSL, INTERFERENCE_COLOR_GLSL])
