"""B200-native MPAS-Atmosphere dycore step (atm_srk3).

Host-side mirror of the reference's time-integration interface plus the
input generators (icosahedral SCVT mesh, Jablonowski-Williamson initial state,
init-time derived mesh fields).  The compute path is the C-ABI CUDA library in
``csrc/`` (``libmpasb.so``); there is no CPU fallback for the step.
"""

__version__ = "0.1.0"
