"""Worker of tests/test_gpu_pinned_path.py::test_nccl_data_parallel_step_two_ranks (launched by torchrun, one rank
per GPU, NCCL).  Prints NCCL_WORKER_OK on success; any assertion fails the launch."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lightweight-multi-modal-scene-understanding-via-knowledge-distillation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from src.data_loading.synthetic_frames import make_frames
    from src.models.camera_encoder import TwinLiteEncoder
    from src.models.fusion_module import CompleteSegmentationModel
    from src.models.lidar_encoder import LiDAREncoder
    from src.training.parallel import allreduce_gradients_
    from src.training.trainer import Trainer

    def make(ft, oc):
        return CompleteSegmentationModel(TwinLiteEncoder(return_multiscale=True), LiDAREncoder("spatial", grid_size=(64, 64)),
                                         num_classes=2, fusion_type=ft, fusion_out_channels=oc,
                                         camera_fpn_stages=["stage3", "stage4", "stage5"], camera_fpn_channels=128,
                                         output_mode="same").to(dev)

    torch.manual_seed(1234 + rank)                      # replicas start DIFFERENT: the Trainer must synchronise them
    student = make("weighted", 128)
    torch.manual_seed(99)                               # the frozen teacher is the same everywhere
    teacher = make("concat", 256)
    tr = Trainer(student, [], [], dev, class_weights=[0.4, 3.5], save_dir=os.path.join(sys.argv[1], f"r{rank}"), teacher=teacher,
                 verbose=False, amp_dtype=torch.bfloat16, use_cuda_graph=True, graph_warmup_steps=1)
    student.train()
    opt = tr.optimizer

    def gathered(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return out

    # 1. construction synchronised parameters, moments and buffers with rank 0
    for t in [opt.flat_param] + [b for b in student.buffers()]:
        parts = gathered(t)
        assert all(torch.equal(parts[0], p) for p in parts[1:]), "replicas differ after Trainer construction"

    # 2. the all-reduced bucket is the sum of the per-rank gradients (different frames per rank)
    b = make_frames(2, 4000, seed=1000 * rank + 7, device=dev)
    opt.detach_grads()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = student(b["image"], b["points"])
    loss = tr.criterion(logits.float(), b["segmentation"])
    loss.backward()
    opt.gather_grads_()
    local_grad = opt.flat_grad.clone()
    parts = gathered(local_grad)
    allreduce_gradients_(opt.flat_grad)
    want = torch.stack(parts).double().sum(0)
    err = ((opt.flat_grad.double() - want).norm() / want.norm()).item()
    assert err < 1e-6, f"all-reduced bucket != sum of per-rank gradients ({err})"
    assert not torch.equal(parts[0], parts[1]), "per-rank gradients should differ (different frames)"
    opt.zero_grad()
    # nothing of that hand-made step may outlive it: a live autograd graph keeps the parameters' AccumulateGrad nodes --
    # and the (legacy default) stream they were created on -- alive, and a later CUDA-graph capture would reuse them
    del logits, loss
    # the running statistics moved differently on each rank in that forward: put the replicas back together
    tr.sync_replicas()

    # 3. eager step, then captured + replayed steps: parameters bit-identical across ranks afterwards
    for i in range(4):
        b = make_frames(2, 4000, seed=1000 * rank + i, device=dev)
        terms, _ = tr.training_step(b["image"], b["points"], b["segmentation"])
        assert torch.isfinite(terms[:4]).all()
    assert len(tr._graphs) == 1
    parts = gathered(opt.flat_param)
    assert all(torch.equal(parts[0], p) for p in parts[1:]), "parameters diverged across ranks"
    moved = (parts[0] - gathered(opt.exp_avg)[0]).abs().max().item()
    assert moved > 0

    # 4. the overlapped two-bucket all-reduce (early bucket launched from the sentinel parameter's hook on a communication
    #    stream, late bucket in line) gives the same parameters as the single in-line bucket
    assert tr._plan and tr._plan["late_span"][0] == 0 and tr._plan["early_span"][1] == opt.flat_grad.numel(), tr._plan
    assert tr._plan["late_span"][1] == tr._plan["early_span"][0] and 0 < tr._plan["late_span"][1] * 4 <= tr.overlap_tail_bytes + 64
    # learning rate 0: the parameters stay put, so the all-reduced gradients of the two trainers must agree step by step to
    # the noise of floating-point atomics (comparing AdamW updates instead would measure AdamW's own chaos: its
    # normalised step turns every noise-level sign flip into a 2 * lr difference); two runs of the SAME path already differ
    # by ~1e-4 here (floating-point atomics, ReLU masks at their threshold), a lost bucket would show as O(1)
    grads = []
    for overlap in (True, False):
        torch.manual_seed(7)
        s2 = make("weighted", 128)
        t2 = Trainer(s2, [], [], dev, lr=0.0, class_weights=[0.4, 3.5], save_dir=os.path.join(sys.argv[1], f"o{rank}{overlap}"),
                     teacher=teacher, verbose=False, amp_dtype=None, use_cuda_graph=False, overlap_allreduce=overlap)
        s2.train()
        per_step = []
        for i in range(3):
            b = make_frames(2, 3000, seed=1000 * rank + 30 + i, device=dev)
            t2.training_step(b["image"], b["points"], b["segmentation"])
            per_step.append(t2.optimizer.flat_grad.clone())
        assert bool(t2._plan) == overlap
        grads.append(per_step)
    for i, (ga, gb) in enumerate(zip(*grads)):
        err = ((ga - gb).double().norm() / gb.double().norm()).item()
        assert gb.abs().max().item() > 0 and err < 2e-3, f"step {i}: overlapped all-reduce changes the reduced gradient: {err}"

    # 5. clean teardown with a captured all-reduce alive until now
    torch.cuda.synchronize()
    dist.barrier()
    tr.release_graphs()
    dist.destroy_process_group()
    print(f"NCCL_WORKER_OK rank {rank}", flush=True)


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        # a failed rank must not leave the other one waiting in a collective (or itself in the communicator's teardown)
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
