// dev_raster.cuh - the analytic (cover/area) rasterizer, restated as a ONE-SCANLINE-AT-A-TIME stepper.
//
// What the reference does (blend2d/raster/analyticrasterizer_p.h):
//   prepare_ref()  :289-360   integer DDA setup for a line p0 -> p1 (24.8 fixed point, p0.y < p1.y)
//   advanceToY()   :371-459   closed-form jump of the DDA state to the start of any scanline
//   rasterize<>()  :466-1165  walks a band of scanlines and adds, per touched cell x of a scanline,
//                             cell[x] += (cover << 9) - area; cell[x + 1] += area          (cell_merge, :1202-1210)
//
// The reference's own unit test (raster/analyticrasterizer_test.cpp:34-157) pins that rasterizing with ANY band
// height (1..32) and jumping with advanceToY() produce identical state; we rely on exactly that invariant: a GPU
// thread prepares an edge, jumps to the first scanline of its tile with edge_advance_to_y(), and then calls
// edge_step_scanline() once per tile row - i.e. the reference algorithm with a band height of one.  The three line
// classes (vertical/single-cell, steep, shallow) keep the reference's integer arithmetic (error terms, rounding
// and update order) but are written as straight-line per-scanline code instead of the banded goto machine.
//
// Cells are u32 and wrap; sums are order independent, so any number of threads may add concurrently.
#pragma once
#include "dev_common.cuh"

namespace b2d {

struct EdgeState {
  int ex0, ey0, ex1, ey1;
  int fx0, fy0, fx1, fy1;
  int x_err, y_err;
  int x_dlt, y_dlt;
  int x_rem, y_rem;
  int x_lift, y_lift;
  int dx, dy;
  int saved_fy1;
  uint32_t flags;
  uint32_t sign_mask;     // 0 or 0xFFFFFFFF (AnalyticRasterizer::_sign_mask)
};

enum : uint32_t {
  kEdgeInitialScanline = 1u,   // AnalyticState::kFlagInitialScanline
  kEdgeVertOrSingle    = 2u,   // AnalyticState::kFlagVertOrSingle
  kEdgeRightToLeft     = 4u    // AnalyticState::kFlagRightToLeft
};

// iter -= step; if (iter < 0) { acc++; iter += correction; }       (AnalyticUtils::acc_err_step, :71-82)
B2D_HD void err_step(int& acc, int& iter, int step, int correction) {
  iter -= step;
  int mask = iter >> 31;
  acc -= mask;
  iter += mask & correction;
}
B2D_HD void err_step_u(uint32_t& acc, int& iter, int step, int correction) {
  iter -= step;
  int mask = iter >> 31;
  acc -= uint32_t(mask);
  iter += mask & correction;
}

// q = a / b, r = a % b for a < 2^64; takes the 32-bit divider whenever the dividend fits (always the case for
// canvases up to 65535 px: coordinates are below 2^24 in 24.8 fixed point).
B2D_HD_COLD uint64_t udiv64_wide(uint64_t a, uint32_t b) { return a / b; }

B2D_HD void udivmod64(uint64_t a, uint32_t b, uint32_t& q, uint32_t& r) {
  if ((a >> 32) == 0) {
    uint32_t a32 = uint32_t(a);
    q = a32 / b;
    r = a32 - q * b;
  }
  else {
    uint64_t q64 = udiv64_wide(a, b);
    q = uint32_t(q64);
    r = uint32_t(a - q64 * b);
  }
}

// `count` steps at once (AnalyticUtils::acc_err_multi_step, :84-100).
B2D_HD void err_multi_step(int& acc, int& iter, int step, int correction, int count) {
  int64_t i = int64_t(uint32_t(iter));
  i -= int64_t(uint64_t(uint32_t(step)) * uint32_t(count));
  if (i < 0) {
    uint64_t num = uint64_t(-i) + uint32_t(correction) - 1u;
    uint32_t q, r;
    udivmod64(num, uint32_t(correction), q, r);
    int n = int(q);
    acc += n;
    i += int64_t(correction) * n;
  }
  iter = int(i);
}

B2D_HD uint32_t apply_sign(uint32_t v, uint32_t sign_mask) { return (v ^ sign_mask) - sign_mask; }

// Returns false for lines that do not cross a scanline boundary in y (p0.y == p1.y).  (prepare_ref, :289-360)
B2D_HD bool edge_prepare(EdgeState& s, int x0, int y0, int x1, int y1, uint32_t sign_bit) {
  if (y0 == y1) return false;

  s.sign_mask = 0u - sign_bit;
  s.dx = x1 - x0;
  s.dy = y1 - y0;
  s.flags = kEdgeInitialScanline;
  if (s.dx < 0) { s.flags |= kEdgeRightToLeft; s.dx = -s.dx; }

  s.ex0 = x0 >> kA8Shift;
  s.ey0 = y0 >> kA8Shift;
  s.ex1 = x1 >> kA8Shift;
  s.ey1 = (y1 - 1) >> kA8Shift;

  s.fx0 = x0 & kA8Mask;
  s.fy0 = y0 & kA8Mask;
  s.fx1 = x1 & kA8Mask;
  s.fy1 = ((y1 - 1) & kA8Mask) + 1;

  s.saved_fy1 = s.fy1;
  if (s.ey0 != s.ey1) s.fy1 = kA8Scale;

  s.x_err = s.y_err = s.x_dlt = s.y_dlt = s.x_rem = s.y_rem = s.x_lift = s.y_lift = 0;

  if (s.ex0 == s.ex1 && (s.ey0 == s.ey1 || s.dx == 0)) {
    s.flags |= kEdgeVertOrSingle;
    return true;
  }

  uint64_t x_base = uint64_t(uint32_t(s.dx)) * kA8Scale;
  uint64_t y_base = uint64_t(uint32_t(s.dy)) * kA8Scale;

  uint32_t q, r;
  udivmod64(x_base, uint32_t(s.dy), q, r); s.x_lift = int(q); s.x_rem = int(r);
  udivmod64(y_base, uint32_t(s.dx), q, r); s.y_lift = int(q); s.y_rem = int(r);

  s.x_dlt = s.dx;
  s.y_dlt = s.dy;
  s.x_err = (s.dy >> 1) - 1;
  s.y_err = (s.dx >> 1) - 1;

  if (s.ey0 != s.ey1) {
    uint64_t p = uint64_t(uint32_t(kA8Scale - s.fy0)) * uint32_t(s.dx);
    udivmod64(p, uint32_t(s.dy), q, r);
    s.x_dlt  = int(q);
    s.x_err -= int(r);
    err_step(s.x_dlt, s.x_err, 0, s.dy);
  }

  if (s.ex0 != s.ex1) {
    uint64_t p = uint64_t((s.flags & kEdgeRightToLeft) ? uint32_t(s.fx0) : uint32_t(kA8Scale - s.fx0)) * uint32_t(s.dy);
    udivmod64(p, uint32_t(s.dx), q, r);
    s.y_dlt  = int(q);
    s.y_err -= int(r);
    err_step(s.y_dlt, s.y_err, 0, s.dx);
  }

  s.y_dlt += s.fy0;
  return true;
}

// Jumps to the beginning of scanline `y_target` (ey0 < y_target <= ey1).  (advanceToY, :371-459)
B2D_HD void edge_advance_to_y(EdgeState& s, int y_target) {
  if (y_target <= s.ey0) return;

  if (!(s.flags & kEdgeVertOrSingle)) {
    int ny = y_target - s.ey0;

    s.x_dlt += s.x_lift * (ny - 1);
    err_multi_step(s.x_dlt, s.x_err, s.x_rem, s.dy, ny - 1);

    if (s.flags & kEdgeRightToLeft) {
      s.fx0 -= s.x_dlt;
      if (s.fx0 < 0) {
        int nx = -(s.fx0 >> kA8Shift);
        s.ex0 -= nx;
        s.fx0 &= kA8Mask;
        err_multi_step(s.y_dlt, s.y_err, s.y_rem, s.dx, nx);
        s.y_dlt += s.y_lift * nx;
      }

      if (!(s.dy >= s.dx)) {
        if (!s.fx0) {
          s.fx0 = kA8Scale;
          s.ex0--;
          err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
          s.y_dlt += s.y_lift;
        }
      }

      if (y_target == s.ey1 && s.dy >= s.dx) {
        s.fy1 = s.saved_fy1;
        s.x_dlt = ((s.ex0 - s.ex1) << kA8Shift) + s.fx0 - s.fx1;
      }
      else {
        s.x_dlt = s.x_lift;
        err_step(s.x_dlt, s.x_err, s.x_rem, s.dy);
      }
    }
    else {
      s.fx0 += s.x_dlt;
      if (s.fx0 >= kA8Scale) {
        int nx = s.fx0 >> kA8Shift;
        s.ex0 += nx;
        s.fx0 &= kA8Mask;
        err_multi_step(s.y_dlt, s.y_err, s.y_rem, s.dx, nx);
        s.y_dlt += s.y_lift * nx;
      }

      if (y_target == s.ey1 && s.dy >= s.dx) {
        s.fy1 = s.saved_fy1;
        s.x_dlt = ((s.ex1 - s.ex0) << kA8Shift) + s.fx1 - s.fx0;
      }
      else {
        s.x_dlt = s.x_lift;
        err_step(s.x_dlt, s.x_err, s.x_rem, s.dy);
      }
    }

    if (s.dy >= s.dx) {
      s.y_dlt &= kA8Mask;
    }
    else {
      int y = ny;
      if (s.flags & kEdgeInitialScanline) y--;
      s.y_dlt -= y * kA8Scale;
    }
  }
  else {
    if (y_target == s.ey1) s.fy1 = s.saved_fy1;
  }

  s.fy0 = 0;
  s.ey0 = y_target;
  s.flags &= ~kEdgeInitialScanline;
}

// Rasterizes scanline `s.ey0` and moves the state to the next one.  Returns true when the edge has ended.
//
// `Sink::merge(x, cover, area)` must perform  cell[x] += (cover << 9) - area;  cell[x + 1] += area.
//
// kWindow (one-row use only - tile_rasterize_edge_row; the state is NOT valid afterwards): the sink only keeps the cells
// of a window of columns [win_lo(), win_hi()), everything left of it as one sum (add_left) and nothing right of it.  A
// shallow edge can cross hundreds of cells in one scanline; instead of walking them
//   * the cells before the window are jumped over with the same closed form advanceToY() uses for whole scanlines
//     (err_multi_step on the y accumulator: the covers of j cells are j * y_lift + the corrections of j steps),
//   * the walk stops at the far side of the window: the covers of a scanline add up to the signed y-extent of the edge
//     in the row, so what is left of the window is that total minus the covers already seen.
template<bool kWindow = false, typename Sink>
B2D_HD bool edge_step_scanline(EdgeState& s, Sink& sink) {
  const uint32_t sm = s.sign_mask;
  s.ey0 += 1;
  const bool more = s.ey0 <= s.ey1;           // scanlines remain after this one

  if (s.flags & kEdgeVertOrSingle) {
    // (:493-571) one cell per scanline, x never changes.
    uint32_t area = uint32_t(s.fx0) + uint32_t(s.fx1);
    uint32_t cover = apply_sign(uint32_t(s.fy1 - s.fy0), sm);
    sink.merge(s.ex0, cover, cover * area);
    if (!more) return true;
    s.fy0 = 0;
    s.fy1 = (s.ey0 == s.ey1) ? s.saved_fy1 : int(kA8Scale);
    return false;
  }

  if (s.dy >= s.dx) {
    // Steep line: at most two cells per scanline (:572-869).
    uint32_t area = uint32_t(s.fx0);
    uint32_t cov;

    if (s.flags & kEdgeRightToLeft) {
      s.fx0 -= s.x_dlt;
      if (s.fx0 < 0) {
        s.ex0--;
        s.fx0 += kA8Scale;
        s.y_dlt &= kA8Mask;

        if (!area) {
          area = kA8Scale;
          err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
          s.y_dlt += s.y_lift;
          cov = apply_sign(uint32_t(s.fy1 - s.fy0), sm);
          sink.merge(s.ex0, cov, cov * (area + uint32_t(s.fx0)));
        }
        else {
          cov = apply_sign(uint32_t(s.y_dlt - s.fy0), sm);
          sink.merge(s.ex0 + 1, cov, cov * area);
          cov = apply_sign(uint32_t(s.fy1 - s.y_dlt), sm);
          sink.merge(s.ex0, cov, cov * (uint32_t(s.fx0) + kA8Scale));
          err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
          s.y_dlt += s.y_lift;
        }
      }
      else {
        cov = apply_sign(uint32_t(s.fy1 - s.fy0), sm);
        sink.merge(s.ex0, cov, cov * (area + uint32_t(s.fx0)));
      }

      s.fy0 = 0;
      if (!more) return true;

      if (s.ey0 == s.ey1) {
        s.fy1 = s.saved_fy1;
        s.x_dlt = ((s.ex0 - s.ex1) << kA8Shift) + s.fx0 - s.fx1;
      }
      else {
        s.x_dlt = s.x_lift;
        err_step(s.x_dlt, s.x_err, s.x_rem, s.dy);
      }
      return false;
    }
    else {
      s.fx0 += s.x_dlt;
      if (s.fx0 <= kA8Scale) {
        cov = apply_sign(uint32_t(s.fy1 - s.fy0), sm);
        sink.merge(s.ex0, cov, cov * (area + uint32_t(s.fx0)));
        if (s.fx0 == kA8Scale) {
          s.ex0++;
          s.fx0 = 0;
          s.y_dlt += s.y_lift;
          err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
        }
      }
      else {
        s.ex0++;
        s.fx0 &= kA8Mask;
        s.y_dlt &= kA8Mask;
        cov = apply_sign(uint32_t(s.y_dlt - s.fy0), sm);
        sink.merge(s.ex0 - 1, cov, cov * (area + kA8Scale));
        cov = apply_sign(uint32_t(s.fy1 - s.y_dlt), sm);
        sink.merge(s.ex0, cov, cov * uint32_t(s.fx0));
        s.y_dlt += s.y_lift;
        err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
      }

      s.fy0 = 0;
      if (!more) return true;

      if (s.ey0 == s.ey1) {
        s.fy1 = s.saved_fy1;
        s.x_dlt = ((s.ex1 - s.ex0) << kA8Shift) + s.fx1 - s.fx0;
      }
      else {
        s.x_dlt = s.x_lift;
        err_step(s.x_dlt, s.x_err, s.x_rem, s.dy);
      }
      return false;
    }
  }

  // Shallow line: a run of cells per scanline (:870-1164).  `x_local` is the 24.8 x position at scanline entry.
  int x_local = (s.ex0 << kA8Shift) + s.fx0;
  uint32_t cover, area;
  bool run;                      // true: walk a multi-cell run; false: the scanline was a single cell

  if (s.flags & kEdgeRightToLeft) {
    if (s.flags & kEdgeInitialScanline) {
      s.flags &= ~kEdgeInitialScanline;
      cover = apply_sign(uint32_t(s.y_dlt - s.fy0), sm);
      run = (s.fx0 - s.x_dlt < 0);
      if (!run) {
        x_local -= s.x_dlt;
        cover = apply_sign(uint32_t(s.fy1 - s.fy0), sm);
        area = cover * uint32_t(s.fx0 * 2 - s.x_dlt);
        sink.merge(s.ex0, cover, area);
        if ((x_local & kA8Mask) == 0) {
          s.y_dlt += s.y_lift;
          err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
        }
        s.x_dlt = s.x_lift;
        err_step(s.x_dlt, s.x_err, s.x_rem, s.dy);
      }
    }
    else {
      if (!more) {
        // Last scanline of the line (:997-1017): exact remaining delta.
        s.x_dlt = x_local - ((s.ex1 << kA8Shift) + s.fx1);
        s.fy1 = s.saved_fy1;
      }
      s.ex0 = (x_local - 1) >> kA8Shift;
      s.fx0 = ((x_local - 1) & kA8Mask) + 1;

      if (!more && s.fx0 - s.x_dlt >= 0) {
        cover = apply_sign(uint32_t(s.fy1), sm);
        area = cover * uint32_t(s.fx0 * 2 - s.x_dlt);
        sink.merge(s.ex0, cover, area);
        return true;
      }

      s.y_dlt -= kA8Scale;
      cover = apply_sign(uint32_t(s.y_dlt), sm);
      run = true;
    }

    if (run) {
      x_local -= s.x_dlt;
      int ex_local = x_local >> kA8Shift;
      int fx_local = x_local & kA8Mask;
      area = cover * uint32_t(s.fx0);

      uint32_t total = 0, done = 0;                     // kWindow: signed y-extent of the edge in this row / covers seen
      if constexpr (kWindow) {
        total = apply_sign(uint32_t(s.fy1 - s.fy0), sm);
        // cells at or right of win_hi() are dropped: jump over them (right to left: they come first)
        const int j = s.ex0 - tmax(ex_local, sink.win_hi() - 1);
        if (j >= 2) {
          int corr = 0;
          err_multi_step(corr, s.y_err, s.y_rem, s.dx, j - 1);
          const uint32_t tn = uint32_t(s.y_lift) * uint32_t(j - 1) + uint32_t(corr);
          s.y_dlt += int(tn);
          done = cover + apply_sign(tn, sm);
          cover = uint32_t(s.y_lift);
          err_step_u(cover, s.y_err, s.y_rem, s.dx);
          s.y_dlt += int(cover);
          cover = apply_sign(cover, sm);
          area = cover * kA8Scale;
          s.ex0 -= j;
        }
      }

      while (s.ex0 != ex_local) {
        if constexpr (kWindow) { if (s.ex0 < sink.win_lo() - 1) { sink.add_left((total - done) << 9); return true; } }   // the rest is left of the window
        sink.merge(s.ex0, cover, area);
        if constexpr (kWindow) done += cover;
        cover = uint32_t(s.y_lift);
        err_step_u(cover, s.y_err, s.y_rem, s.dx);
        s.y_dlt += int(cover);
        cover = apply_sign(cover, sm);
        area = cover * kA8Scale;
        s.ex0--;
      }

      cover += apply_sign(uint32_t(s.fy1 - s.y_dlt), sm);
      area = cover * (uint32_t(fx_local) + kA8Scale);
      sink.merge(s.ex0, cover, area);

      if (fx_local == 0) {
        s.y_dlt += s.y_lift;
        err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
      }
      s.x_dlt = s.x_lift;
      err_step(s.x_dlt, s.x_err, s.x_rem, s.dy);
    }

    s.fy0 = 0;
    s.fy1 = kA8Scale;
    s.ex0 = (x_local - 1) >> kA8Shift;
    s.fx0 = ((x_local - 1) & kA8Mask) + 1;
    return !more;
  }
  else {
    if (s.flags & kEdgeInitialScanline) {
      s.flags &= ~kEdgeInitialScanline;
      cover = apply_sign(uint32_t(s.y_dlt - s.fy0), sm);
      run = (s.fx0 + s.x_dlt > kA8Scale);
      if (!run) {
        x_local += s.x_dlt;
        cover = apply_sign(uint32_t(s.fy1 - s.fy0), sm);
        area = cover * (uint32_t(s.fx0) * 2 + uint32_t(s.x_dlt));
        sink.merge(s.ex0, cover, area);
        if (s.fx0 + s.x_dlt == kA8Scale) {
          s.y_dlt += s.y_lift;
          err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
        }
        s.x_dlt = s.x_lift;
        err_step(s.x_dlt, s.x_err, s.x_rem, s.dy);
      }
    }
    else {
      if (!more) {
        s.x_dlt = ((s.ex1 << kA8Shift) + s.fx1) - x_local;
        s.fy1 = s.saved_fy1;
      }
      s.ex0 = x_local >> kA8Shift;
      s.fx0 = x_local & kA8Mask;

      if (!more && s.fx0 + s.x_dlt <= kA8Scale) {
        cover = apply_sign(uint32_t(s.fy1), sm);
        area = cover * (uint32_t(s.fx0) * 2 + uint32_t(s.x_dlt));
        sink.merge(s.ex0, cover, area);
        return true;
      }

      s.y_dlt -= kA8Scale;
      cover = apply_sign(uint32_t(s.y_dlt), sm);
      run = true;
    }

    if (run) {
      x_local += s.x_dlt;
      int ex_local = (x_local - 1) >> kA8Shift;
      int fx_local = ((x_local - 1) & kA8Mask) + 1;
      area = cover * (uint32_t(s.fx0) + kA8Scale);

      if constexpr (kWindow) {
        // cells up to win_lo() - 2 only feed the sum left of the window: jump over them (left to right: they come first)
        const int j = tmin(ex_local, sink.win_lo() - 1) - s.ex0;
        if (j >= 2) {
          int corr = 0;
          err_multi_step(corr, s.y_err, s.y_rem, s.dx, j - 1);
          const uint32_t tn = uint32_t(s.y_lift) * uint32_t(j - 1) + uint32_t(corr);
          s.y_dlt += int(tn);
          sink.add_left((cover + apply_sign(tn, sm)) << 9);
          cover = uint32_t(s.y_lift);
          err_step_u(cover, s.y_err, s.y_rem, s.dx);
          s.y_dlt += int(cover);
          cover = apply_sign(cover, sm);
          area = cover * kA8Scale;
          s.ex0 += j;
        }
      }

      while (s.ex0 != ex_local) {
        if constexpr (kWindow) { if (s.ex0 >= sink.win_hi()) return true; }   // the rest is right of the window: dropped
        sink.merge(s.ex0, cover, area);
        cover = uint32_t(s.y_lift);
        err_step_u(cover, s.y_err, s.y_rem, s.dx);
        s.y_dlt += int(cover);
        cover = apply_sign(cover, sm);
        area = cover * kA8Scale;
        s.ex0++;
      }

      cover += apply_sign(uint32_t(s.fy1 - s.y_dlt), sm);
      area = cover * uint32_t(fx_local);
      sink.merge(s.ex0, cover, area);

      if (fx_local == kA8Scale) {
        s.y_dlt += s.y_lift;
        err_step(s.y_dlt, s.y_err, s.y_rem, s.dx);
      }
      s.x_dlt = s.x_lift;
      err_step(s.x_dlt, s.x_err, s.x_rem, s.dy);
    }

    s.fy0 = 0;
    s.fy1 = kA8Scale;
    s.ex0 = x_local >> kA8Shift;
    s.fx0 = x_local & kA8Mask;
    return !more;
  }
}

} // namespace b2d
