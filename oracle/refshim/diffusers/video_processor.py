import numpy as np
import PIL.Image
import torch
import torch.nn.functional as F


class VideoProcessor:
    def __init__(self, vae_scale_factor=8, do_resize=True, do_normalize=True):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image, height=None, width=None):
        if isinstance(image, PIL.Image.Image):
            image = [image]
        if isinstance(image, (list, tuple)) and isinstance(image[0], PIL.Image.Image):
            arrs = [np.asarray(im.convert("RGB").resize((width, height), PIL.Image.LANCZOS), dtype=np.float32) / 255.0 for im in image]
            t = torch.from_numpy(np.stack(arrs)).permute(0, 3, 1, 2)
        elif torch.is_tensor(image):
            t = image if image.ndim == 4 else image[None]
            if height and width and tuple(t.shape[-2:]) != (height, width):
                t = F.interpolate(t, size=(height, width))
            if t.min() < 0:  # already in [-1, 1]: diffusers warns and skips the normalisation
                return t
        else:
            raise ValueError(type(image))
        return 2.0 * t - 1.0

    def postprocess_video(self, video, output_type="np"):
        v = (video / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 4, 1).float()
        return v.cpu().numpy() if output_type == "np" else v
