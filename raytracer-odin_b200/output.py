"""Mirror of output.odin: get_rgb_image (mode Mean, output.odin:30-80) and save_result (:82-107).
Odin keeps this stage in production; the harness needs it to look at images."""
import numpy as np


def tone_mapping_aces(x):  # output.odin:21-28
    x = x.astype(np.float32)
    a, b, c, d, e = (np.float32(v) for v in (2.51, 0.03, 2.43, 0.59, 0.14))
    return np.clip((x * (a * x + b)) / (x * (c * x + d) + e), 0, 1).astype(np.float32)


def get_rgb_image(stats: np.ndarray, width: int, height: int) -> np.ndarray:
    with np.errstate(invalid="ignore", divide="ignore"):
        raw = stats["total"] / stats["count"].astype(np.float32)[:, None]
    raw = np.maximum(raw, np.float32(0))
    g = np.power(tone_mapping_aces(raw), np.float32(1 / 2.2))
    rgb = np.floor(g * np.float32(255) + np.float32(0.5))  # linalg.round for non-negative values
    return np.nan_to_num(rgb, nan=0).astype(np.uint8).reshape(height, width, 3)


def save_result(stats: np.ndarray, width: int, height: int, file_path: str):
    rgb = get_rgb_image(stats, width, height)
    if file_path.endswith(".ppm"):  # output.odin:88-94
        with open(file_path, "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (width, height))
            f.write(rgb.tobytes())
    elif file_path.endswith(".png"):
        import cv2

        cv2.imwrite(file_path, rgb[:, :, ::-1])
    else:
        raise RuntimeError(f"Unsupported file format: {file_path}")  # output.odin:105
