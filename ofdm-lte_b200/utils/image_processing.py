"""Image <-> bit payload adapters used by the GUIs and the image tests of the reference
(utils/image_processing.py:9-255).  Host-side I/O only: the bit vectors made here are what simulate_* and
lte_b200.sweep.payload_sweep carry.  Bit order is MSB first per byte, row-major RGB -- the order
`lte_bits_to_indices` consumes."""
import numpy as np

_RGB = 'RGB'
_PEAK = 255.0


def _image_module():
    from PIL import Image              # imported lazily: the engine itself never needs PIL
    return Image


def _open_rgb(path):
    im = _image_module().open(path)
    return im if im.mode == _RGB else im.convert(_RGB)


def _pixels(obj):
    """PIL image or array-like -> ndarray."""
    return np.array(obj) if isinstance(obj, _image_module().Image) else np.asarray(obj)


def _same_shape(ref, other):
    """Resize `other` to the height / width of `ref` when they differ (the reference's behaviour)."""
    if ref.shape == other.shape:
        return other
    resized = _image_module().fromarray(other).resize((ref.shape[1], ref.shape[0]))
    return np.array(resized)


def _psnr(mse):
    return float('inf') if mse == 0 else 20.0 * np.log10(_PEAK / np.sqrt(mse))


def _mse(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.mean(d * d))


class ImageProcessor:
    """Static helpers with the reference's names and return conventions."""

    @staticmethod
    def load_image_pil(image_path):
        return _open_rgb(image_path)

    @staticmethod
    def image_to_bits(image_path):
        """-> (uint8 bit vector of H*W*3*8 entries, {'height', 'width', 'channels', 'dtype'})."""
        px = np.array(_open_rgb(image_path))
        rows, cols, planes = px.shape
        meta = dict(height=rows, width=cols, channels=planes, dtype=str(px.dtype))
        return np.unpackbits(px.reshape(-1)), meta

    @staticmethod
    def bits_to_image(bits, metadata):
        """Zero-pad or cut the stream to the image size and rebuild the picture; a black image of the right
        size if the stream cannot be reshaped."""
        Image = _image_module()
        shape = (metadata['height'], metadata['width'], metadata['channels'])
        want = int(np.prod(shape)) * 8
        stream = np.zeros(want, dtype=np.uint8)
        have = np.asarray(bits).astype(np.uint8)[:want]
        stream[:len(have)] = have
        try:
            return Image.fromarray(np.packbits(stream).reshape(shape), _RGB)
        except Exception as exc:
            print(f"Error al reconstruir imagen: {exc}")
            return Image.new(_RGB, (shape[1], shape[0]), color='black')

    @staticmethod
    def calculate_psnr(original_img, reconstructed_img):
        ref = _pixels(original_img)
        return _psnr(_mse(ref, _same_shape(ref, _pixels(reconstructed_img))))

    @staticmethod
    def calculate_psnr_bits(original_bits, reconstructed_bits):
        """PSNR between the byte streams the two bit vectors pack to, over their common prefix."""
        n = min(len(original_bits), len(reconstructed_bits))
        packed = [np.packbits(np.asarray(v[:n]).astype(np.uint8)) for v in (original_bits, reconstructed_bits)]
        return _psnr(_mse(packed[0], packed[1]))          # packbits zero-fills the last partial byte on both sides

    @staticmethod
    def calculate_ssim(original_img, reconstructed_img):
        try:
            from skimage.metrics import structural_similarity
        except ImportError:
            print("scikit-image no disponible para calcular SSIM")
            return None
        ref = _pixels(original_img)
        return structural_similarity(ref, _same_shape(ref, _pixels(reconstructed_img)), channel_axis=2, data_range=255)

    @staticmethod
    def save_comparison(original_path, reconstructed_img, output_path):
        """Original and reconstruction side by side."""
        Image = _image_module()
        left = Image.open(original_path)
        right = reconstructed_img if reconstructed_img.size == left.size else reconstructed_img.resize(left.size)
        canvas = Image.new(_RGB, (2 * left.size[0], left.size[1]))
        for k, im in enumerate((left, right)):
            canvas.paste(im, (k * left.size[0], 0))
        canvas.save(output_path)
        return canvas
