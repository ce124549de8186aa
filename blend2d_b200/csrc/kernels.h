// kernels.h - parameter blocks and launch entry points of the libb2dgpu kernels (see kernels.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/b2dgpu.h"

namespace b2d {

struct BuildParams {
  const double* vertices;                 // x,y pairs
  const b2dgpu_segment* segments;
  uint32_t segment_count;
  const b2dgpu_command* commands;
  const b2dgpu_geometry_state* states;
  uint32_t* seg_counts;                   // count pass: edges per segment
  const uint32_t* seg_offsets;            // write pass: exclusive scan of seg_counts (segment_count + 1 entries)
  b2dgpu_edge* edges;                     // device edge array (caller-supplied edges first, built edges after)
  uint32_t edge_base;                     // index of the first built edge
  uint32_t edge_capacity;                 // total capacity of `edges`
  int4* cmd_bbox_fixed;                   // per command: min x, min y, max x, max y in 24.8
  uint32_t* error_flag;
};

struct FinalizeParams {
  const b2dgpu_command* commands;
  uint32_t command_count;
  const uint32_t* seg_offsets;
  uint32_t edge_base;
  const int4* cmd_bbox_fixed;
  int4* cmd_bbox_px;                      // clipped pixel box [x0,x1) x [y0,y1) used for tile culling
  uint2* cmd_edges;                       // (first edge, edge count)
  int width;
  int y_begin, y_end;                     // rows of the full image owned by this target
};

struct TileParams {
  uint8_t* dst;                           // first byte of row `y_begin`
  intptr_t dst_stride;
  int tiles_x, tiles_y;
  int y_begin;
  const b2dgpu_command* commands;
  uint32_t command_count;
  const int4* cmd_bbox_px;
  const uint2* cmd_edges;
  // Per-band command lists (k_bin_*): band b owns cells [band_off[b], band_off[b + 1]) in submission order; a cell is
  // (command index, (min cell x, ~max cell x) of what the command can touch in that band).  bin_state[1] == 0: the lists
  // did not fit their buffer and were not built - the compositor then scans every command (bounding boxes only).
  const uint32_t* band_off;
  const uint32_t* cell_cmd;
  const uint2* cell_ext;
  const uint32_t* bin_state;              // [0] cells needed, [1] lists valid, [2] per-cell edge lists valid; nullptr = no lists
  // Per (band, command) cell: the edges of the command that cross the band, as indices into `edges`:
  // band_edges[cell_edge_off[cell] .. cell_edge_off[cell + 1]).  Phase 1 of the compositor walks these instead of all
  // the edges of the command (the reference's per-band edge lists, edgestorage_p.h:38-178).
  const uint32_t* cell_edge_off;
  const uint32_t* band_edges;
  const b2dgpu_edge* edges;
  const b2dgpu_fetch_data* fetch_data;
  const uint8_t* bayer;
  int origin_x, origin_y;
  unsigned long long* pixel_counter;
};

// K0: glyph instancing (dev_glyph.cuh).  One thread per instance writes the glyph's vertices and path segments into the
// batch's arrays, behind the uploaded ones.
struct GlyphParams {
  const uint32_t* cache;                  // device mirror of the caller's glyph cache
  const b2dgpu_glyph_instance* instances;
  uint32_t instance_count;
  double* vertices;                       // the batch's vertex array (x, y pairs)
  b2dgpu_segment* segments;               // the batch's segment array
  uint32_t* error_flag;
};
int launch_glyph_instances(const GlyphParams& P, cudaStream_t s);

// Gradient tables built on the device: one warp per request.
struct LutParams {
  const b2dgpu_lut_request* requests;
  uint32_t request_count;
  const b2dgpu_gradient_stop* stops;
  const uint32_t* table_offsets;          // per request: word offset of its table inside `tables`
  uint32_t* tables;
};
int launch_build_luts(const LutParams& P, cudaStream_t s);

// Binning (K1d): per-band ordered command lists with x-extents, the GPU form of the reference's per-band edge lists
// (raster/edgestorage_p.h:38-178) and of its band-by-band command walk (raster/workerproc.cpp:166-255).
struct BinParams {
  const b2dgpu_command* commands;
  uint32_t command_count;
  const int4* cmd_bbox_px;
  const uint2* cmd_edges;
  const b2dgpu_edge* edges;
  int y_begin, tile_h, tiles_y;
  uint32_t* cm_count;                     // per command: bands its pixel box covers           (command_count + 1)
  uint32_t* cm_base;                      // exclusive scan of cm_count                        (command_count + 1)
  uint32_t* band_count;                   // difference array of the commands per band         (tiles_y + 2)
  uint32_t* band_prefix;                  // its exclusive scan: commands of band b = [b + 1]  (tiles_y + 2)
  uint32_t* band_off;                     // exclusive scan of band_count                      (tiles_y + 1)
  uint32_t* cm_index;                     // [cm_base[c] + band - first band of c] -> cell     (capacity)
  uint32_t* cell_cmd;                     // band-major cells                                  (capacity)
  uint2* cell_ext;                        //                                                   (capacity)
  uint32_t capacity;
  uint32_t* state;                        // [0] cells needed, [1] lists valid, [2] edge lists valid, [3] (edge, band) pairs needed
  uint32_t* scan_scratch;
  uint32_t* cell_edge_cnt;                // (edge, band) pairs per cell; counted down to zero again by k_bin_edges     (capacity)
  uint32_t* cell_edge_off;                // exclusive scan of it                                                      (capacity + 1)
  uint32_t* band_edges;                   // edge indices, grouped by cell                                             (edge_list_capacity)
  uint32_t edge_list_capacity;
  uint32_t* scan_scratch2;                // for the scan over the cells
};
size_t bin_scratch_items(uint32_t command_count, int tiles_y);
size_t bin_cell_scan_scratch_items(uint32_t capacity);

// One solid box fill over a large region (fill_all / clear_all / big FillRectA): pure streaming, see k_stream_solid.
struct SolidStreamParams {
  uint8_t* dst;                           // first byte of row `y_begin` of the target
  intptr_t dst_stride;
  int x0w, x1w;                           // 32-bit words per row to touch: pixels for 32-bpp targets, 4-pixel groups for A8
  int y0, y1;                             // rows, relative to the first row held by the target
  uint32_t mode;                          // 0 SrcOver, 1 SrcCopy with m < 255, 2 SrcCopy with m == 255 (store only)
  uint32_t src;                           // premultiplied source pixel (A8: alpha replicated into the four bytes)
  uint32_t mask;                          // 1..255
  unsigned long long pixels;              // pixels the fill composites (added to the counter by one thread)
  unsigned long long* pixel_counter;
};

// One FillBoxA with a gradient / pattern source over a large region of a 32-bit target: see stream.cu (k_stream_one).
struct StreamOneParams {
  uint8_t* dst;                           // first byte of row `y_begin`
  intptr_t dst_stride;
  int y_begin;
  int x0, y0, x1, y1;                     // the box in pixels, clipped to the target (y absolute)
  uint32_t fetch_type, src_format, comp_op, alpha;
  uint32_t stage_lut;                     // entries of the gradient table if it is to be staged in shared memory by one
                                          // cp.async.bulk (stream.cu); 0 = look it up in global memory
  const b2dgpu_fetch_data* fd;            // device copy of the command's FetchData
  const uint8_t* bayer;
  int origin_x, origin_y;
  unsigned long long pixels;
  unsigned long long* pixel_counter;
};

// Each launcher returns the number of kernels it launched.
int launch_count_edges(const BuildParams& P, cudaStream_t s);
int launch_write_edges(const BuildParams& P, cudaStream_t s);
size_t scan_scratch_items(uint32_t n);
// `n_dev` (optional): the item count lives on the device; at most `n` items, out[min(n, *n_dev)] = total.
int launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch, uint32_t* total_out, cudaStream_t s, const uint32_t* n_dev = nullptr);
int launch_init_bbox(int4* bbox, uint32_t n, cudaStream_t s);
int launch_analytic_bbox(const b2dgpu_command* cmds, uint32_t ncmd, const b2dgpu_edge* edges, int4* bbox, cudaStream_t s);
int launch_finalize_commands(const FinalizeParams& P, cudaStream_t s);
int choose_tile_height(int tiles_x, int rows, int sm_count);
int launch_binning(const BinParams& B, cudaStream_t s);
int launch_tile_render(const TileParams& P, int bpp, int tile_h, cudaStream_t s);
int launch_box_stream(const TileParams& P, int bpp, const int* box, int sm_count, cudaStream_t s);
int launch_stream_solid(const SolidStreamParams& P, int sm_count, cudaStream_t s);
int launch_stream_one(const StreamOneParams& P, int sm_count, cudaStream_t s);

} // namespace b2d
