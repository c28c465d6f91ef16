"""examples/sphere.rs of the reference, through the B200 engine: volumes of the unit balls.

    1D: 2 * 1
    2D: 2 * integral(sqrt(1 - x^2), -1..1)                         (QAGS, default config)
    3D: 2 * integral2(sqrt(1 - x^2 - y^2), DynamicY(-1, 1, circle))
The closures become registry functors (semicircle_p, hemisphere_p with p0 = 1; y-range 'circle').
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "gkquad-rs_b200")]
from gkquad_b200 import double, single  # noqa: E402


def main():
    v1 = 2.0 * 1.0
    v2 = 2.0 * single.integral(single.DeviceIntegrand("semicircle_p", [1.0]), (-1.0, 1.0)).unwrap().estimate
    rng = double.DynamicY.new(-1.0, 1.0, "circle", [1.0])
    v3 = 2.0 * double.integral2(double.DeviceIntegrand2("hemisphere_p", [1.0]), rng).unwrap().estimate
    print("1D:", v1)
    print("2D:", v2)
    print("3D:", v3)
    return v1, v2, v3


if __name__ == "__main__":
    main()
