"""Generate tests/golden/*.npz from the reference's own test fixtures.

Run HERE (the build container), where /root/reference exists; the GPU box never reads
/root/reference, it only loads the committed .npz files.

Inputs (read-only, reference test/test.cpp:41-44):
    test/Pterois_volitans_Manado-e_edit_smallest.h   input JPEG (256x185 gray)
    test/edges.h, test/linesDark.h, test/linesBright.h   expected-output JPEGs
The `xxd -i` byte arrays are parsed as text, the JPEG bytes decoded with cv2
(IMREAD_GRAYSCALE, as test/test.cpp:53-56 does) and the DECODED PIXELS are stored,
so the fixtures do not depend on the libjpeg build of whichever box runs the tests.

Also stores float-level oracle outputs on the fish image (cv2 4.13.0) as regression
vectors for the oracle itself and as GPU parity fixtures.
"""
import os
import re
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import cvsteer_ref as ref  # noqa: E402

REF = "/root/reference/test"


def read_xxd(path):
    txt = open(path).read()
    data = bytes(int(h, 16) for h in re.findall(r"0x([0-9a-fA-F]{2})", txt))
    n = int(re.search(r"_len\s*=\s*(\d+)\s*;", txt).group(1))
    assert len(data) == n, (path, len(data), n)
    return data


def decode(path):
    img = cv2.imdecode(np.frombuffer(read_xxd(path), np.uint8), cv2.IMREAD_GRAYSCALE)
    assert img is not None and img.ndim == 2
    return img


def main():
    fish = decode(f"{REF}/Pterois_volitans_Manado-e_edit_smallest.h")
    edges = decode(f"{REF}/edges.h")
    dark = decode(f"{REF}/linesDark.h")
    bright = decode(f"{REF}/linesBright.h")
    assert fish.shape == (185, 256) and int(fish.sum()) == 6968201, (fish.shape, fish.sum())
    np.savez_compressed(os.path.join(HERE, "fish_fixture.npz"), fish=fish, edges_gt=edges,
                        lines_dark_gt=dark, lines_bright_gt=bright)

    # float-level oracle outputs (cv2 4.13.0) on the fish image
    f2, (g2, h2, e, mag, phase) = ref.g2_full(fish)
    out = {k: getattr(f2, k) for k in ref.SteerableFiltersG2.PLANES}
    out.update(c1=f2.c1, c2=f2.c2, c3=f2.c3, theta=f2.theta, strength=f2.strength,
               g2=g2, h2=h2, e=e, magnitude=mag, phase=phase)
    f4 = ref.SteerableFiltersG4(fish)
    out.update({k: getattr(f4, k) for k in ref.SteerableFiltersG4.PLANES})
    g4s, h4s = f4.steer_scalar(0.3)
    g4m, h4m, mag4, ph4 = f4.steer_map_full(f2.theta)
    out.update(g4_s03=g4s, h4_s03=h4s, g4_map=g4m, h4_map=h4m, mag4=mag4, phase4=ph4)
    lv = ref.pyramid(fish, 5)
    for i, l in enumerate(lv[1:], 1):
        out[f"pyr{i}"] = l
    out = {k: np.ascontiguousarray(v, np.float32) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "fish_oracle_cv2_4_13.npz"), **out)
    taps = {n: getattr(f2, n)[0] for n in ("g1", "g2", "g3", "h1", "h2", "h3", "h4")}
    taps.update({"G4_" + n: getattr(f4, n)[0] for n in
                 ("g1", "g2", "g3", "g4", "g5", "h1", "h2", "h3", "h4", "h5", "h6")})
    np.savez(os.path.join(HERE, "taps_default.npz"), **taps)
    print("wrote fixtures; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
