"""diffusers.utils names used by the reference scripts."""
USE_PEFT_BACKEND = False


def is_accelerate_available():
    return False


def load_image(image):
    """diffusers.utils.load_image for local paths / PIL images (no network)."""
    from PIL import Image, ImageOps

    if isinstance(image, str):
        image = Image.open(image)
    image = ImageOps.exif_transpose(image)
    return image.convert("RGB")
