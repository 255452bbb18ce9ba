"""Mirror of reference models/base_model.py:81-103 (BaseModel.__call__, evaluation path)."""


class BaseModel:
    def __init__(self, model, opt):
        self.model = model
        self.opt = opt
        self.netD = None  # the discriminator only ranks num_samples > 1 candidates

    def __call__(self, batch, isval=False, num_steps=1, return_batch=False):
        if not isval:
            raise NotImplementedError("the GAN training step (base_model.py:105-148) is out of scope")
        t_losses, output_images = self.model(batch, self.netD)
        if getattr(self.opt, "normalize_image", False):
            for k in output_images.keys():
                if "Img" in k:
                    output_images[k] = 0.5 * output_images[k] + 0.5
        if return_batch:
            return t_losses, output_images, batch
        return t_losses, output_images
