"""Image <-> bit payload adapters of the GUIs and image tests (reference utils/image_processing.py:9-255).
Pure host-side I/O (PIL + NumPy): the bits produced here are what simulate_* / payload_sweep carry,
so the existing front-ends can drive the GPU engine unchanged.  Bit order is np.unpackbits /
np.packbits (MSB first), the same order `lte_bits_to_indices` consumes."""
import numpy as np


def _pil():
    from PIL import Image
    return Image


class ImageProcessor:
    @staticmethod
    def image_to_bits(image_path):
        """-> (bits uint8 [H*W*3*8], metadata dict) (reference :12-48)."""
        img = _pil().open(image_path)
        if img.mode != 'RGB':
            img = img.convert('RGB')
        arr = np.array(img)
        h, w, c = arr.shape
        return np.unpackbits(arr.flatten()), {'height': h, 'width': w, 'channels': c, 'dtype': str(arr.dtype)}

    @staticmethod
    def bits_to_image(bits, metadata):
        """Truncate / zero-pad to the image size and rebuild it (reference :50-89)."""
        Image = _pil()
        h, w, c = metadata['height'], metadata['width'], metadata['channels']
        need = h * w * c * 8
        bits = np.asarray(bits)
        bits = np.pad(bits, (0, need - len(bits)), 'constant') if len(bits) < need else bits[:need]
        try:
            return Image.fromarray(np.packbits(bits.astype(np.uint8)).reshape(h, w, c).astype(np.uint8), 'RGB')
        except Exception as e:      # the reference returns a black image of the right size
            print(f"Error al reconstruir imagen: {e}")
            return Image.new('RGB', (w, h), color='black')

    @staticmethod
    def calculate_psnr(original_img, reconstructed_img):
        Image = _pil()
        a = np.array(original_img) if isinstance(original_img, Image.Image) else np.asarray(original_img)
        b = np.array(reconstructed_img) if isinstance(reconstructed_img, Image.Image) else np.asarray(reconstructed_img)
        if a.shape != b.shape:
            b = np.array(Image.fromarray(b).resize((a.shape[1], a.shape[0])))
        mse = np.mean((a.astype(float) - b.astype(float)) ** 2)
        return float('inf') if mse == 0 else 20 * np.log10(255.0 / np.sqrt(mse))

    @staticmethod
    def calculate_psnr_bits(original_bits, reconstructed_bits):
        """PSNR of the byte streams the two bit arrays pack to (reference :131-168)."""
        n = min(len(original_bits), len(reconstructed_bits))
        a, b = np.asarray(original_bits[:n]), np.asarray(reconstructed_bits[:n])
        pad = (8 - n % 8) % 8
        if pad:
            a = np.concatenate([a, np.zeros(pad, dtype=int)])
            b = np.concatenate([b, np.zeros(pad, dtype=int)])
        mse = np.mean((np.packbits(a.astype(np.uint8)).astype(float) - np.packbits(b.astype(np.uint8)).astype(float)) ** 2)
        return float('inf') if mse == 0 else 20 * np.log10(255.0 / np.sqrt(mse))

    @staticmethod
    def calculate_ssim(original_img, reconstructed_img):
        try:
            from skimage.metrics import structural_similarity as ssim
        except ImportError:
            print("scikit-image no disponible para calcular SSIM")
            return None
        Image = _pil()
        a = np.array(original_img) if isinstance(original_img, Image.Image) else np.asarray(original_img)
        b = np.array(reconstructed_img) if isinstance(reconstructed_img, Image.Image) else np.asarray(reconstructed_img)
        if a.shape != b.shape:
            b = np.array(Image.fromarray(b).resize((a.shape[1], a.shape[0])))
        return ssim(a, b, channel_axis=2, data_range=255)

    @staticmethod
    def save_comparison(original_path, reconstructed_img, output_path):
        Image = _pil()
        orig = Image.open(original_path)
        if orig.size != reconstructed_img.size:
            reconstructed_img = reconstructed_img.resize(orig.size)
        w, h = orig.size
        comp = Image.new('RGB', (w * 2, h))
        comp.paste(orig, (0, 0))
        comp.paste(reconstructed_img, (w, 0))
        comp.save(output_path)
        return comp

    @staticmethod
    def load_image_pil(image_path):
        img = _pil().open(image_path)
        return img.convert('RGB') if img.mode != 'RGB' else img
