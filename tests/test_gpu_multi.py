"""2-GPU tests (-m gpu; skipped on a one-GPU box -- run them with `gpurun --gpus 2`): the multi-GPU layer end to end,
one PROCESS per GPU, through (i) the C++ host mirror (`BSMR-sddmm -r rank -w world -u idfile`, NCCL id exchanged
through a file) and (ii) the Python mirror (multigpu.ShardedSDDMM over sddmm_mgpu_*; id exchanged with
torch.distributed).  Every rank ends with the whole P after sddmm_mgpu_gather; it is checked against the oracle."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from cases import ROOT, gen

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")


def _nccl_lib():
    try:
        import nvidia.nccl as n
        p = os.path.join(list(n.__path__)[0], "lib", "libnccl.so.2")
        return p if os.path.exists(p) else ""
    except Exception:
        return ""


@needs2
def test_cli_two_ranks_file_rendezvous(tmp_path):
    exe = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")
    assert os.access(exe, os.X_OK), "CLI not built"
    S = gen.rmat(12, 8, 4)
    mtx = str(tmp_path / "g.mtx")
    gen.write_mtx(mtx, S, order="col")
    idf = str(tmp_path / "nccl.id")
    procs = []
    for r in range(2):
        env = dict(os.environ, LOCAL_RANK=str(r))
        if _nccl_lib():
            env["SDDMM_B200_NCCL_LIB"] = _nccl_lib()
        procs.append(subprocess.Popen([exe, "-f", mtx, "-k", "64", "-a", "0.3", "-d", "0.3", "-b", "16", "-c", "1", "-r", str(r),
                                       "-w", "2", "-u", idf], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env))
    outs = [p.communicate(timeout=600) for p in procs]
    shard = []
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, err[-500:]
        assert "Pass! Result validates successfully." in out, out[-600:]  # every rank holds the whole, correct P
        kv = dict(l.strip("[]").split(" : ", 1) for l in out.replace("], [", "]\n[").splitlines() if " : " in l and l.startswith("["))
        assert int(kv["b200_rank"]) == r and int(kv["b200_world"]) == 2
        shard.append((int(kv["b200_shard_nnz"]), [int(x) for x in kv["b200_shard_panels"].split()]))
    assert shard[0][0] + shard[1][0] == S.nnz
    assert shard[0][1][1] == shard[1][1][0] and shard[0][1][0] == 0            # contiguous panel ranges
    assert abs(shard[0][0] - shard[1][0]) < 0.2 * S.nnz                        # balanced by nnz


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from __graft_entry__ import load_package
        pkg = load_package()
        from sddmm_gpu_b200 import multigpu as mg
        from oracle import oracle as O
        S = pkg.generators.rmat(12, 8, 4)
        K = 64
        A, B = pkg.generators.dense_operands(S.M, S.N, K)
        g = mg.MultiGpu(rank, world, mg.exchange_unique_id(device="cuda"))
        ro = torch.from_numpy(S.row_off.view(np.int32)).cuda()
        ci = torch.from_numpy(S.col_idx.view(np.int32)).cuda()
        sh = mg.ShardedSDDMM(g, ro, ci, S.M, S.N, alpha=0.3, delta=0.3, block_size=16)
        rr = O.row_reorder(S, 0.3, 16)["reorderedRows"]
        same_order = bool(np.array_equal(sh.R.cpu().numpy().view(np.uint32), rr))  # rank 0's order reached everyone
        dA = torch.from_numpy(A).cuda()
        dB = torch.from_numpy(B).cuda() if rank == 0 else torch.zeros((S.N, K), device="cuda")
        sh.replicate_B(dB)
        P = torch.zeros(S.nnz, device="cuda")
        sh.run(dA, dB, P)
        torch.cuda.synchronize()
        mine = int((P != 0).sum())
        g.gather(P)
        torch.cuda.synchronize()
        Pref = O.sddmm_cpu(S, A, B)
        ok = O.check_data(Pref, P.cpu().numpy()) == 0
        # host-buffer pass of the whole job: 1/world slices over PCIe, all-gather over NVLink, P reduced onto rank 0
        hP = np.full(S.nnz, np.nan, np.float32) if rank == 0 else None
        for _ in range(2):  # twice: the staging buffers are reused and must not accumulate
            g.run_host(sh.layout, A, B, hP, 0)
        if rank == 0:
            ok = ok and O.check_data(Pref, hP) == 0
        # the same with page-locked buffers: every rank reads only its share of the REFERENCED rows (half of an R-MAT
        # graph's rows and columns are empty) through the mapped pointers; packed all-gather, unpack on the device
        full_bytes = g.host_traffic()
        pA, pB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
        hP2 = np.full(S.nnz, np.nan, np.float32) if rank == 0 else None
        for _ in range(2):
            g.run_host(sh.layout, pA.numpy(), pB.numpy(), hP2, 0)
        ok = ok and 0 < g.host_traffic() < 0.8 * full_bytes
        if rank == 0:
            ok = ok and O.check_data(Pref, hP2) == 0
        # cost calibration: the cuts may move, the union must still cover every non-zero exactly once
        sh.calibrate(dA, dB, P, rounds=2, passes=2)
        P.zero_()
        sh.run(dA, dB, P)
        g.gather(P)
        torch.cuda.synchronize()
        ok = ok and O.check_data(Pref, P.cpu().numpy()) == 0
        q.put((rank, ok and same_order, sh.my_nnz, mine, [int(c) for c in sh.cuts]))
        g.close()
    finally:
        dist.destroy_process_group()


@needs2
@pytest.mark.timeout(600)
def test_python_two_ranks_sharded_sddmm_nccl():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    S = gen.rmat(12, 8, 4)
    assert all(r[1] for r in res)
    assert res[0][4] == res[1][4]
    assert res[0][2] + res[1][2] == S.nnz
