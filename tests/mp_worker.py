"""Worker of tests/test_multigpu.py, launched once per rank by torch.distributed.run.

mode "oracle": every rank owns one block in a CPU oracle instance; halo exchanges are done on host
               arrays over gloo (mpas_model_b200.multigpu.HostExchanger).  No GPU needed.
mode "gpu":    every rank owns one block on its own GPU; exchanges go through the library
               (pack kernel -> NCCL send/recv -> unpack kernel).
Rank 0 gathers the owned elements of every rank and writes them to <out>.npz."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    mode, n_cells, n_lev, n_scal, n_steps, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    from mpas_model_b200 import multigpu as mg
    dist = mg.init_distributed("gloo")
    if mode == "gpu":
        g, block, cfg, ex, _ = mg.setup_rank(n_cells, n_lev, n_scal, rank, world, device=int(os.environ.get("LOCAL_RANK", rank)))
        dt = cfg["config_dt"]
        for _ in range(n_steps):
            g.atm_srk3(dt)
            g.mpas_pool_shift_time_levels()
    else:
        from oracle.oracle import OracleDycore
        box = [mg.prepare_blocks(n_cells, n_lev, n_scal, world) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        rec = mg.load_block(box[0], rank)
        block, cfg, ex = rec["block"], rec["cfg"], rec["ex"]
        cfg = dict(cfg, **json.loads(os.environ.get("MPASB_TEST_CFG", "{}")))      # namelist overrides of the test
        g = OracleDycore(block, cfg, rank=rank)
        hx = mg.HostExchanger(dist, rank, ex)
        xch = lambda group: hx.exchange(g, group)
        dt = cfg["config_dt"]
        xch("initialization:u")
        g.atm_init_coupled_diagnostics()
        g.atm_init_solve_diagnostics(dt)
        xch("initialization:pv_edge,ru,rw")
        for _ in range(n_steps):
            mg.srk3_host_exchange(g, cfg, dt, xch)
            g.mpas_pool_shift_time_levels()
    glob = mg.gather_owned(dist, g, block)
    if rank == 0:
        np.savez(out, **glob)
    dist.barrier()


if __name__ == "__main__":
    main()
