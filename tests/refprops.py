"""Defaults the reference's init() functions give each geometrictransform element (struct field names),
needed because the oracle/_ref harness starts from a zeroed element struct. Citations: the DEFAULT_*
macros of gst/geometrictransform/gst<element>.c and gstcirclegeometrictransform.c:73-75."""
import math

CIRCLE = {"x_center": 0.5, "y_center": 0.5, "radius": 0.35}
CIRCLE_ELEMENTS = ("bulge", "circle", "kaleidoscope", "pinch", "sphere", "twirl", "waterripple", "stretch", "tunnel")
DEFAULTS = {
    "fisheye": {}, "tunnel": {},
    "bulge": {"zoom": 3.0},
    "circle": {"angle": 0.0, "spread_angle": math.pi, "height": 20},
    "kaleidoscope": {"angle": 0.0, "angle2": 0.0, "sides": 3},
    "pinch": {"intensity": 0.5},
    "rotate": {"angle": 0.0},
    "sphere": {"refraction": 1.5},
    "twirl": {"angle": math.pi},
    "waterripple": {"amplitude": 10.0, "phase": 0.0, "wavelength": 16.0},
    "stretch": {"intensity": 0.5},
    "square": {"width": 0.5, "height": 0.5, "zoom": 2.0},
    "mirror": {"mode": 0},
    "perspective": {"matrix_0": 1, "matrix_1": 0, "matrix_2": 0, "matrix_3": 0, "matrix_4": 1, "matrix_5": 0,
                    "matrix_6": 0, "matrix_7": 0, "matrix_8": 1},
}
# the element's init(): most subclasses select clamp, the base default is ignore
DEFAULT_OFF_EDGE = {"circle": "ignore", "rotate": "ignore", "perspective": "ignore"}

# property sets exercised by the parity tests (struct field names)
CASES = {
    "fisheye": [{}],
    "bulge": [{}, {"zoom": 7.5, "x_center": 0.3, "radius": 0.6}],
    "circle": [{}, {"angle": 1.1, "spread_angle": 2.0, "height": 33, "y_center": 0.7}],
    "kaleidoscope": [{}, {"angle": 0.7, "angle2": -0.3, "sides": 5, "radius": 0.0}, {"sides": 7, "x_center": 0.25}],
    "pinch": [{}, {"intensity": -0.8, "radius": 0.9}],
    "rotate": [{}, {"angle": 0.6}],
    "sphere": [{}, {"refraction": 2.2, "radius": 0.8}],
    "twirl": [{}, {"angle": -2.0}],
    "waterripple": [{}, {"amplitude": 25.0, "phase": 1.0, "wavelength": 9.0, "radius": 0.7}],
    "stretch": [{}, {"intensity": 0.9}],
    "tunnel": [{}, {"radius": 0.2, "x_center": 0.4}],
    "square": [{}, {"width": 0.3, "height": 0.8, "zoom": 4.0}],
    "mirror": [{"mode": m} for m in range(4)],
    "perspective": [{}, {"matrix_0": 1.1, "matrix_1": 0.1, "matrix_2": -5, "matrix_3": 0.05, "matrix_4": 0.9, "matrix_5": 3,
                         "matrix_6": 0.0002, "matrix_7": 0.0001, "matrix_8": 1.0}],
}


def full(element, props):
    d = dict(DEFAULTS[element])
    if element in CIRCLE_ELEMENTS:
        d.update(CIRCLE)
    d.update(props)
    return d
