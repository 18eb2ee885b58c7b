import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import __graft_entry__ as g
pkg = g.load_package()
from sph_b200 import slab, scenes
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
pos, mass, params, dt = scenes.dam_break_scene(0.004 / world ** (1/3))
n = len(pos); nsr = float(params["neighbor_search_radius"])
cells = slab.axis_cells(pos, 2, nsr); cuts = slab.plan_cuts(cells, world, 2)
store = slab.GpuStore(pkg, int(1.6 * n / world) + 400000, local, params, stream=torch.cuda.current_stream().cuda_stream)
r = slab.SlabRank(store, rank, cuts, 2, 2, n, pos.min(0), pos.max(0), n // world)
r.load_initial(pos, None, mass, nsr)
dev = store.device
def sync(): torch.cuda.synchronize()
T = {k: 0.0 for k in ("pack", "gather", "a2a", "append", "step")}
store.ctx.set_option(pkg.capi.OPT_STAGE_TIMING, 1)
for it in range(13):
    sync(); dist.barrier(); t0 = time.perf_counter()
    counts = r.pack(); sync(); t1 = time.perf_counter()
    mine = torch.from_numpy(counts).to(dev)
    table = torch.empty((world, 2 * world), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(table.view(-1), mine)
    send, recv = slab._splits(table.cpu().numpy(), rank); t2 = time.perf_counter()
    n_out, n_in = sum(send), sum(recv)
    dist.all_to_all_single(r.recv[:n_in].view(-1), r.send[:n_out].view(-1), output_split_sizes=[c * 8 for c in recv], input_split_sizes=[c * 8 for c in send])
    sync(); t3 = time.perf_counter()
    store.append(r.recv, n_in, None); sync(); t4 = time.perf_counter()
    store.step(dt); sync(); t5 = time.perf_counter()
    if it == 2:
        store.ctx.reset_stats()
    if it >= 3:
        for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)): T[k] += v
if rank in (0, world // 2):
    st = store.ctx.stats()
    print(rank, {k: round(1e3 * v / 10, 3) for k, v in T.items()}, "n_in", n_in, "size", store.size,
          {k: round(1e3 * st[k] / st["steps"], 3) for k in ("neighbor_search_time", "density_computation_time", "force_computation_time", "integration_time")}, flush=True)
dist.destroy_process_group()
