"""train -- drop-in for /root/reference/tinyfaces/trainer.py:68-90 (the step loop), plus the data-parallel
pieces the reference does not have: one process per GPU, batch-sharded, gradients SUM-reduced over NCCL
(the reference loss is a sum, loss.py:87-88, so the reduction must not divide).
"""
import torch
import torch.distributed as dist


def print_state(idx, epoch, size, loss_cls, loss_reg):
    head = "Epoch: [{0}][{1}/{2}]\t".format(epoch, idx, size) if epoch >= 0 else "Val: [{0}/{1}]\t".format(idx, size)
    print(head + "\tloss_cls: {0:.6f}\tloss_reg: {1:.6f}".format(loss_cls, loss_reg))


def allreduce_gradients(parameters, group=None):
    """SUM all-reduce of every existing .grad as ONE flat fp32 buffer (110.9 MB for the trunk + heads)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def train_step(model, loss_fn, optimizer, img, class_map, regression_map, group=None):
    """One forward / loss / backward / (all-reduce) / SGD step on tensors already on the device."""
    output = model(img)
    loss = loss_fn(output, class_map, regression_map)
    optimizer.zero_grad()
    loss.backward()
    allreduce_gradients([p for g in optimizer.param_groups for p in g["params"]], group)
    optimizer.step()
    return loss


def train(model, loss_fn, optimizer, dataloader, epoch, device):
    """trainer.py:68-90."""
    model = model.to(device)
    model.train()
    for idx, (img, class_map, regression_map) in enumerate(dataloader):
        x = img.float().to(device)
        class_map_var = class_map.float().to(device)
        regression_map_var = regression_map.float().to(device)
        train_step(model, loss_fn, optimizer, x, class_map_var, regression_map_var)
        print_state(idx, epoch, len(dataloader), loss_fn.class_average.average, loss_fn.reg_average.average)
