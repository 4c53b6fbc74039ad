import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as sm
rank=int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev=torch.device("cuda",int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl",device_id=dev)
try:
    t=sm.empty(1<<20,dtype=torch.float32,device=dev)
    h=sm.rendezvous(t,dist.group.WORLD)
    print(rank,"ok world",h.world_size,"multicast_ptr",hex(h.multicast_ptr) if h.multicast_ptr else h.multicast_ptr,"buffers",[hex(p) for p in h.buffer_ptrs][:3],"signal pads",len(h.signal_pad_ptrs), flush=True)
    t.fill_(rank+1); h.barrier()
    if h.multicast_ptr:
        # sanity: one_shot_all_reduce through torch's multimem op
        out=torch.ops.symm_mem.multimem_all_reduce_(t,"sum",dist.group.WORLD.group_name)
        torch.cuda.synchronize(); print(rank,"multimem all_reduce value",float(t[0]), flush=True)
except Exception as e:
    import traceback; traceback.print_exc(); print(rank,"FAILED",repr(e)[:300], flush=True)
dist.barrier(); dist.destroy_process_group()
