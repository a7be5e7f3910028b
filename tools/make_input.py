"""Write an ARTES input tree (input/<name>/artes.in + atmosphere.fits) for one of the synthetic configurations.

usage: python tools/make_input.py c1|c2|c3|c4|c5|lambert [name] [key=value ...]

The reference's own setup tools (python/atmosphere.py, opacity*.py) are Python 2 + astropy and do not run in
this image; the files written here follow their formats so that the host driver
reads exactly what the reference would."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import atmospheres as A


def write_input(atm, name, root=".", **over):
    d = os.path.join(root, "input", name)
    os.makedirs(d, exist_ok=True)
    kw = dict(A.TEMPLATE_ARTES_IN)
    kw.update(atm.artes_in)
    kw.update(over)
    with open(os.path.join(d, "artes.in"), "w") as f:
        f.write("=" * 70 + "\n* ARTES input parameters\n\n" + "-" * 70 + "\n")
        for k, v in kw.items():
            f.write(f"{k}={v}\n")
    A.write_atmosphere_fits(atm, os.path.join(d, "atmosphere.fits"))
    return d


if __name__ == "__main__":
    cfg = sys.argv[1]
    name = sys.argv[2] if len(sys.argv) > 2 and "=" not in sys.argv[2] else cfg
    over = dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a)
    atm = A.lambert_sphere() if cfg == "lambert" else A.CONFIGS[cfg]()
    print(write_input(atm, name, **over))
