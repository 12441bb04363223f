"""Worker of tests/test_ddp_gpu.py::test_recipe_loop_under_ddp: `python -m torch.distributed.run --nproc-per-node 2 tests/ddp_recipe_worker.py <tmp>`.
Runs slam_llm.pipeline.finetune.main (enable_ddp=true: NCCL process group, trainables broadcast from rank 0, per-step ASYNC all-reduce of the
flat gradient arena with the optimizer step deferred behind the next front end, DistributedSampler shards) on a tiny recipe and checks that
both replicas end with identical trainables and that they differ from the initial ones."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "src"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    tmp = sys.argv[1]
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    import slam_llm  # noqa: F401
    from recipe_util import make_data, make_llm_dir, run_config
    from slam_llm.pipeline import finetune
    llm_dir, data_dir, out = os.path.join(tmp, "llm"), os.path.join(tmp, "data"), os.path.join(tmp, f"out{rank}")
    if rank == 0:
        make_llm_dir(llm_dir)
        make_data(data_dir, n=8)
        open(os.path.join(tmp, "ready"), "w").write("1")
    else:
        import time
        while not os.path.exists(os.path.join(tmp, "ready")):
            time.sleep(0.2)
    os.makedirs(out, exist_ok=True)
    cfg = run_config(llm_dir, os.path.join(data_dir, "data.jsonl"), out, os.path.join(ROOT, "tests", "recipe_model.py") + ":model_factory",
                     os.path.join(ROOT, "src/slam_llm/datasets/speech_dataset.py") + ":get_speech_dataset", enable_ddp=True, num_epochs=2,
                     run_validation=False, save_model=False, batch_size_training=2)
    captured = {}
    orig_train = finetune.train

    def spy(model, *a, **kw):
        captured["model"] = model
        captured["before"] = model.b200.arena.param.detach().clone()
        assert model.b200.defer_update and model.ddp_world_size == 2
        return orig_train(model, *a, **kw)
    finetune.train = spy
    results = finetune.main(cfg)
    model = captured["model"]
    model.b200.flush_update()
    p = model.b200.arena.param.detach().clone()
    both = [torch.empty_like(p) for _ in range(2)]
    dist.all_gather(both, p)
    assert torch.equal(both[0], both[1]), "replicas diverged"
    moved = (p - captured["before"]).abs().max().item()
    assert moved > 0, "parameters did not move"
    assert float(results["avg_train_loss"]) > 0
    if rank == 0:
        print(f"DDP_RECIPE_OK moved={moved:.3e} loss={float(results['avg_train_loss']):.4f}", flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
