"""Import shim: the reference imports h5py at module scope (chain.py:3) and h5py / libhdf5 are not installed in
this image.  TEST INFRASTRUCTURE ONLY.

The shim hands the UNMODIFIED reference the package's own HDF5 implementation (bipymc_b200/h5lite.py, loaded by
file path: the reference's processes never import the bipymc_b200 package), so its checkpoint code --
McmcChain.write_chain_h5 / read_chain_h5 (chain.py:59-93), DeMcMpi.save_state / load_state (demc.py:198-233) --
runs as written: `h5py.File(name, "w")`, `create_dataset(..., compression="gzip")`, `del h5f[name]`,
`isinstance(x, h5py.File)`, `h5f[name][:]`.  tests/test_reference_h5_interop.py exchanges checkpoint files between
the reference and bipymc_b200 through it.  (No `version` attribute: bipymc_b200.h5lite.get_h5() tells the shim from
the real library by that.)"""
import importlib.util
import os

_path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "bipymc_b200",
                     "h5lite.py")
_spec = importlib.util.spec_from_file_location("_bipymc_b200_h5lite_for_reference", _path)
_h5lite = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_h5lite)

File = _h5lite.File
Group = _h5lite.Group
Dataset = _h5lite.Dataset
