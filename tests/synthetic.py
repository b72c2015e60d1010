"""Seeded generator of synthetic raster origami systems (BASELINE.json config 5): a
scaffold of n_rows * width domains laid out as a straight line along +x (identities -1..-N, all
orientations (0,0,1), no staples: valid because constraints only apply to bound pairs), and one two-domain
staple type per column and pair of adjacent raster rows. Written in the reference's system-JSON format; no
sequences (use hybridization_pot=Uniform)."""
import json


def raster_system(width, n_rows, cyclic=False):
    assert n_rows % 2 == 0
    n = width * n_rows

    def index(row, col):
        return row * width + (col if row % 2 == 0 else width - 1 - col)

    identities = [[-(i + 1) for i in range(n)]]
    for m in range(n_rows // 2):
        for c in range(width):
            d1, d2 = index(2 * m, c), index(2 * m + 1, c)
            identities.append([d1 + 1, d2 + 1])
    if cyclic:
        # closed rectangle: out along y = 0, back along y = 1 (first and last domains adjacent)
        half = n // 2
        positions = [[i, 0, 0] for i in range(half)] + [[half - 1 - i, 1, 0] for i in range(n - half)]
    else:
        positions = [[i, 0, 0] for i in range(n)]
    chain = {"index": 0, "identity": 0, "positions": positions, "orientations": [[0, 0, 1] for _ in range(n)]}
    return {"origami": {"identities": identities, "cyclic": cyclic,
                        "configurations": [{"step": 0, "chains": [chain]}]}}


def write_raster_system(path, width, n_rows, cyclic=False):
    with open(path, "w") as f:
        json.dump(raster_system(width, n_rows, cyclic), f)
    return path


# Uniform-potential parameters of config 5 (kb K and kb; order of magnitude of 8-bp NN values)
UNIFORM_OPTIONS = {
    "hybridization_pot": "Uniform", "binding_h": -3.0e4, "binding_s": -83, "misbinding_h": -5.0e3, "misbinding_s": -20,
    "stacking_ene": -1000, "domain_type": "ThreeQuarterTurn",
}
