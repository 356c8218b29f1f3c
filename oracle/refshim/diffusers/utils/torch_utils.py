import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """CPU generators draw on the CPU and the result moves to ``device``; a list of generators draws one sample each."""
    device = torch.device(device or "cpu")
    batch = shape[0]
    if isinstance(generator, list) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, list):
        one = (1,) + tuple(shape[1:])
        return torch.cat([randn_tensor(one, generator[i], device, dtype) for i in range(batch)], dim=0)
    rand_device = device
    if generator is not None and generator.device.type != device.type:
        if generator.device.type != "cpu":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {generator.device.type}.")
        rand_device = torch.device("cpu")
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)
