"""TEST INFRASTRUCTURE ONLY (build container, where /root/reference exists): the UNMODIFIED reference classes with the
three touch points of INTEGRATION.md section B applied in a subclass and the ctypes stub of that section executed as
written, bound to the emulator build of the C ABI; compared with the reference's own numpy `calc` on the same files
through the reference's own loaders.

    python tests/emu/reference_binding_check.py <libedk_emu.so>
"""
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

from make_golden import import_reference  # noqa: E402
from oracle import elemental_oracle as orc  # noqa: E402


def main(lib_path):
    lattice = import_reference()
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# lattice/generator/_edk\.py.*?)```", text, re.S).group(1)
    stub = {}
    exec(compile(block.replace('C.CDLL("libedk_sm100a.so")', f"C.CDLL({lib_path!r})"), "INTEGRATION.md", "exec"), stub)
    EdkHandle, MODE = stub["EdkHandle"], stub["EDK_MODE_DERIVATIVE"]

    class BoundElementalGenerator(lattice.ElementalGenerator):
        """lattice/generator/elemental.py with the three touch points of INTEGRATION.md B."""

        def __init__(self, latt_size, gauge_field, eigenvector, num_nabla=0, momentum_list=[(0, 0, 0)]):
            super().__init__(latt_size, gauge_field, eigenvector, num_nabla, momentum_list)
            self._edk = EdkHandle(latt_size, self.Ne, MODE, num_nabla, momentum_list)            # __init__, after :59

        def load(self, key):
            super().load(key)
            self._U_file = np.ascontiguousarray(self.gauge_field.load(key)[:])                     # load, :103

        def calc(self, t):
            eigenvector, V, VPV = self._eigenvector_data, self._V, self._VPV
            for e in range(V.shape[0]):                                                            # unchanged staging loop
                V[e] = eigenvector[t, e]
            self._edk.calc(self._U_file[t], V, VPV)                                                # calc, :297-338
            return VPV

    latt_size, Ne, num_nabla = [4, 2, 2, 2], 4, 2
    Lx, Ly, Lz, Lt = latt_size
    moms = [(0, 0, 0), (0, 0, 1), (1, 1, 0)]
    U = np.stack([orc.synthetic_links(latt_size, t) for t in range(Lt)])
    V = np.stack([orc.synthetic_eigvecs(latt_size, Ne, t) for t in range(Lt)])
    worst = 0.0
    with tempfile.TemporaryDirectory() as tmp:
        prefix = tmp + "/"
        U.astype("<c16").tofile(prefix + "cfg.dat")
        np.save(prefix + "cfg.eigenvector.npy", V.astype("<c16"))
        gauge = lattice.preset.GaugeFieldBinary(prefix, ".dat", [Lt, Lz, Ly, Lx, 4, 3, 3], "<c16")
        evec = lattice.EigenvectorNpy(prefix, ".eigenvector.npy", [Lt, Ne, Lz, Ly, Lx, 3], Ne)
        stock = lattice.ElementalGenerator(latt_size, gauge, evec, num_nabla, moms)
        bound = BoundElementalGenerator(latt_size, gauge, evec, num_nabla, moms)
        stock.load("cfg")
        bound.load("cfg")
        for t in range(Lt):
            ref = np.array(stock.calc(t), copy=True)
            got = bound.calc(t)
            assert got is bound._VPV  # the reference's ownership rule: its own buffer is returned
            scale = np.sqrt((np.abs(ref) ** 2).sum(axis=(-1, -2)))
            floor = 1e-4 * scale.max()
            for d in range(ref.shape[0]):
                for p in range(ref.shape[1]):
                    worst = max(worst, float(np.linalg.norm(got[d, p] - ref[d, p]) / max(scale[d, p], floor)))
    # the elemental files this package writes, read back by the REFERENCE's own readers (lattice/preset.py:129-137,173-181)
    import easydistillation_b200.preset as ours

    rng = np.random.default_rng(1)
    Nop, Nmom, Lt2 = 3, 2, 4
    data = (rng.standard_normal((Nop, Nmom, Lt2, Ne, Ne)) + 1j * rng.standard_normal((Nop, Nmom, Lt2, Ne, Ne))).astype("<c16")
    with tempfile.TemporaryDirectory() as tmp:
        prefix = tmp + "/"
        mm = ours.ElementalBinary(prefix, ".meson", [Nop, Nmom, Lt2, Ne, Ne], Ne).create("cfg", data.shape)
        mm[...] = data
        mm.flush()
        ref_bin = lattice.preset.ElementalBinary(prefix, ".meson", [Nop, Nmom, Lt2, Ne, Ne], Ne).load("cfg")
        assert np.array_equal(ref_bin[2, 1, 3], data[2, 1, 3]) and np.array_equal(ref_bin[0, 0], data[0, 0])
        mm = ours.ElementalNpy(prefix, ".meson.npy").create("cfg", data.shape)
        mm[...] = data
        mm.flush()
        ref_npy = lattice.ElementalNpy(prefix, ".meson.npy", [Nop, Nmom, Lt2, Ne, Ne], Ne).load("cfg")
        assert np.array_equal(ref_npy[1, 0, 2], data[1, 0, 2])
    print(f"EDK_BINDING_OK worst block error {worst:.3e}; elemental files readable by the reference's ElementalBinary / ElementalNpy")
    return 0 if worst < 1e-10 else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
