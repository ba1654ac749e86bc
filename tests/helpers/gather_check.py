"""Run under torchrun with N >= 2 GPUs (tests/test_gpu_multi.py): the batch-sharded native loop + the in-stream all-gather of the
C ABI (dv_comm_all_gather) must equal (a) the torch.distributed gather and (b) the unsharded single-GPU result, bit for bit."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import diffvits_b200
    from diffvits_b200 import dist as dvd
    from diffvits_b200 import synthetic as inputs
    from diffvits_b200.synthetic import UNET_KW
    diffvits_b200.install_dropin()
    from unet1d.unet_1d_condition import UNet1DConditionModel
    torch.manual_seed(1234)
    unet = UNet1DConditionModel(**UNET_KW).eval().to(dev)
    B, T, Tp = 2 * world + 1, 72, 40                       # ragged shards: the remainder goes to the first ranks
    full_in = inputs.make_inputs(B, T, Tp, seed=3)
    x, cond, ehs = (t.to(dev) for t in dvd.shard_inputs(full_in, rank, world))
    betas = inputs.betas_linear()
    # the hoisted text-time embedding is a PyTorch module: evaluate it on the whole batch everywhere and shard the result, so
    # that the comparison below is about the native path only (cuBLAS may pick another algorithm for another batch size)
    with torch.no_grad():
        aug_full = unet.add_embedding(full_in[2].to(dev))
    aug = dvd.shard_inputs((aug_full,), rank, world)[0]
    native = diffvits_b200.NativeDenoiser(unet, cond, ehs, aug=aug)
    lat = diffvits_b200.sample_native(unet, x, cond, ehs, sampler="dpmsolver++", steps=6, betas=betas, native=native)
    via_torch = dvd.gather_latents(lat, B).clone()
    assert not dvd.native_comm_active()
    assert dvd.init_native_comm(dev), "dv_comm_init failed"
    via_abi = dvd.gather_latents(lat, B)
    torch.cuda.synchronize(dev)
    assert via_abi.shape == (B, 100, T) and torch.equal(via_abi, via_torch)
    if rank == 0:
        xa, ca, ea = (t.to(dev) for t in full_in)
        nat_all = diffvits_b200.NativeDenoiser(unet, ca, ea, aug=aug_full)
        whole = diffvits_b200.sample_native(unet, xa, ca, ea, sampler="dpmsolver++", steps=6, betas=betas, native=nat_all)
        assert torch.equal(whole, via_abi), float((whole - via_abi).abs().max())
        print(f"GATHER_OK world={world} B={B}")
    dist.barrier()
    dvd.close_native_comm()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
