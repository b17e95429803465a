"""torchrun worker: lookahead parallelism over NCCL through the plugin surface must reproduce the single-GPU ids."""
import os
import random
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    import lade
    from helpers import build_hf_llama, load_cases
    from lookaheaddecoding_b200 import LookaheadEngine
    from test_gpu_e2e import assert_same_or_tie

    os.environ["USE_LADE"] = "1"
    lade.augment_all()
    ok = True
    for name in ("tiny_bf16_w15n5g15_pool", "tiny_bf16_w5n3g3"):
        c = load_cases()[name]
        model, w = build_hf_llama(c["model"], c["weight_seed"], device=f"cuda:{local}")
        model.generation_config.pad_token_id = 0
        model.generation_config.eos_token_id = None
        # single-GPU engine first (same model object, no process group involved)
        single = LookaheadEngine(model, c["W"], c["N"], c["G"], pool_from_prompt=c["pool_from_prompt"],
                                 max_total_len=len(c["prompt"]) + 64)
        ref = single.generate(c["prompt"], 64, rng=random.Random(5))
        single.close()
        lade.config_lade(LEVEL=c["N"], WINDOW_SIZE=c["W"], GUESS_SET_SIZE=c["G"], DEBUG=1,
                         POOL_FROM_PROMPT=c["pool_from_prompt"], DIST_WORKERS=world, backend="nccl")
        assert lade.distributed() and lade.get_device() == local
        ids = torch.tensor([c["prompt"]], device=f"cuda:{local}")
        random.seed(5 + 17 * rank)       # ranks draw different windows; rank 0's is broadcast
        out = model.generate(ids, attention_mask=torch.ones_like(ids), max_new_tokens=64, do_sample=False)[0].tolist()
        gathered = [None] * world
        dist.all_gather_object(gathered, out)
        assert all(g == gathered[0] for g in gathered), "ranks disagree"
        assert_same_or_tie(out, ref, c["model"], w, f"LP x{world} {name}")
        from lookaheaddecoding_b200.decoding import CONFIG_MAP
        from lookaheaddecoding_b200.decoding import get_engine
        lp_eng = get_engine(model)
        exact = out == ref
        # the same sequence again through the torch all-gather fallback (exchange outside the graph): same ids
        if rank == 0:
            print(f"{name}: LP x{world} log {CONFIG_MAP['log'][-1]} ; single-GPU steps {single.last_steps} ; "
                  f"exchange in library (ncclAllGather in the step graph): {lp_eng.lp_in_library} ; "
                  f"ids identical to the single-GPU run: {exact}")
        assert lp_eng.lp_in_library, "the in-library NCCL exchange was expected on a multi-GPU NCCL run"
    dist.barrier()
    if rank == 0 and ok:
        print("LP_WORKER_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
