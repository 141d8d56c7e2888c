"""The bins' charts: `plot_moving_average` of the reference (src/utils.rs:97-157) — one 600x400
PNG per title with a line per run, the reference's colours, a legend box — drawn with Pillow (the reference uses the
`plotters` crate; pixels differ, the content and the file names do not).  Host-side post-processing of the gathered
per-episode metrics; nothing here touches the device.
"""
import math
import os

# plotters' BLUE GREEN CYAN RED YELLOW MAGENTA, then the bins' own darker shades (src/bin/taxi.rs:112-123)
COLORS = [(0, 0, 255), (0, 255, 0), (0, 255, 255), (255, 0, 0), (255, 255, 0), (255, 0, 255), (150, 0, 0), (0, 0, 150),
          (0, 150, 0), (50, 0, 0), (0, 0, 50), (0, 50, 0)]
TITLES = (("train_rewards", "Train Rewards"), ("train_episodes_length", "Train Episodes Length"), ("train_errors", "Training Error"),
          ("test_rewards", "Test Rewards"), ("test_episodes_length", "Test Episodes Length"))   # src/bin/taxi.rs:205-223
SIZE = (600, 400)
LEFT = BOTTOM = 40          # set_label_area_size(Left / Bottom, 40), utils.rs:133-134


def value_range(values):
    """utils.rs:106-130: the longest series' length and the min / max over all series with `f64::min` / `f64::max`
    (which SKIP NaNs), then the guard — a flat or NaN range becomes [-1, 1]."""
    def fmin(a, b):
        return b if math.isnan(a) else (a if math.isnan(b) else min(a, b))

    def fmax(a, b):
        return b if math.isnan(a) else (a if math.isnan(b) else max(a, b))
    def reduce(fn, v):
        acc = float(v[0])
        for x in v[1:]:
            acc = fn(acc, float(x))
        return acc
    max_len = len(values[0])
    lo, hi = reduce(fmin, values[0]), reduce(fmax, values[0])
    for v in values[1:]:
        max_len = max(max_len, len(v))
        lo_i, hi_i = reduce(fmin, v), reduce(fmax, v)
        if lo_i < lo:           # a NaN on either side compares false: an all-NaN FIRST series poisons the range (-> guard)
            lo = lo_i
        if hi_i > hi:
            hi = hi_i
    if lo == hi or math.isnan(lo) or math.isnan(hi):
        lo, hi = -1.0, 1.0
    return max_len, lo, hi


def plot_moving_average(values, colors, legends, title, out_dir=".", verbose=True):
    """Writes `<out_dir>/<title>.png` and returns its path.  `values` is a list of series (lists of f64)."""
    try:
        from PIL import Image, ImageDraw, ImageFont
    except ImportError as exc:   # pragma: no cover
        raise RuntimeError("charts need Pillow (PIL); the curves are also available as JSON (driver --out)") from exc
    max_len, lo, hi = value_range(values)
    if verbose:                                               # utils.rs:124-125
        print("max len %d" % max_len)
        print("%s | %s\n" % (repr(lo), repr(hi)))
    img = Image.new("RGB", SIZE, (255, 255, 255))
    d = ImageDraw.Draw(img)
    font = ImageFont.load_default()
    w, h = SIZE
    top = 46                                                   # caption(title, ("sans-serif", 40))
    x0, y0, x1, y1 = LEFT, top, w - 10, h - BOTTOM
    d.text((w // 2, 6), title, fill=(0, 0, 0), font=font, anchor="ma")
    n_x = max(1, max_len)

    def px(i, y):
        fx = x0 + (x1 - x0) * (i / n_x)
        fy = y1 - (y1 - y0) * ((y - lo) / (hi - lo))
        return fx, fy
    for k in range(11):                                        # configure_mesh(): light grid, ticks and their labels
        gx = x0 + (x1 - x0) * k / 10.0
        gy = y1 - (y1 - y0) * k / 10.0
        d.line([(gx, y0), (gx, y1)], fill=(225, 225, 225))
        d.line([(x0, gy), (x1, gy)], fill=(225, 225, 225))
        d.text((gx, y1 + 4), "%d" % round(n_x * k / 10.0), fill=(0, 0, 0), font=font, anchor="ma")
        d.text((x0 - 3, gy), "%.3g" % (lo + (hi - lo) * k / 10.0), fill=(0, 0, 0), font=font, anchor="rm")
    d.line([(x0, y0), (x0, y1), (x1, y1)], fill=(0, 0, 0))
    for v, c in zip(values, colors):                           # LineSeries per run; NaN points break the line
        run = []
        for i, y in enumerate(v):
            y = float(y)
            if math.isnan(y) or math.isinf(y):
                if len(run) > 1:
                    d.line(run, fill=tuple(c))
                run = []
            else:
                run.append(px(i, min(max(y, lo), hi)))
        if len(run) > 1:
            d.line(run, fill=tuple(c))
        elif len(run) == 1:
            d.point(run, fill=tuple(c))
    # configure_series_labels(): a bordered, nearly opaque box with a 20-pixel sample of each line
    lh = 11
    bw = 30 + max(int(d.textlength(s, font=font)) for s in legends) + 8
    bx1, by0 = x1 - 4, y0 + 4
    bx0, by1 = bx1 - bw, by0 + lh * len(legends) + 6
    d.rectangle([bx0, by0, bx1, by1], fill=(250, 250, 250), outline=(0, 0, 0))
    for r, (s, c) in enumerate(zip(legends, colors)):
        yy = by0 + 4 + r * lh + lh // 2
        d.line([(bx0 + 4, yy), (bx0 + 24, yy)], fill=tuple(c))
        d.text((bx0 + 28, yy), s, fill=(0, 0, 0), font=font, anchor="lm")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "%s.png" % title)
    img.save(path)
    return path


def plot_experiment(result, out_dir=".", verbose=True):
    """The five charts at the end of every bin's main() (src/bin/taxi.rs:205-223) from run_experiment()'s dict."""
    legends = [s.replace("ε", "eps") for s in result["legends"]]   # the default bitmap font is Latin-1
    return [plot_moving_average(result[key], COLORS[:len(result[key])], legends[:len(result[key])], title, out_dir, verbose)
            for key, title in TITLES]
