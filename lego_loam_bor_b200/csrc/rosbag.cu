// rosbag.cu -- ROS-less reader of rosbag v2.0 files for the bag mode of the reference's main loop
// (LeGO-LOAM/src/main.cpp:26-35 bag.open, :60-76 rosbag::View over the lidar topic + instantiate<sensor_msgs::PointCloud2>
// + IP.cloudHandler).  SURVEY.md section 8 f4.  Host code only (it lives in a .cu file so that the one library build picks
// it up).  The bag format and the ROS message serialisation are restated from their published specifications (rosbag
// "Bag format 2.0", sensor_msgs/PointCloud2.msg); ROS is not vendored in the reference tree: parity unpinned.
//
// A bag is "#ROSBAG V2.0\n" followed by records; a record is <u32 header_len><header><u32 data_len><data>, a header is a
// list of <u32 field_len><name>=<value> fields, and the `op` field tells the record type: 0x03 bag header, 0x05 chunk (its
// data is a run of records again), 0x07 connection (topic + a second header in the data with type / md5sum /
// message_definition), 0x02 message data (conn, time; data = the serialised message), 0x04 index data, 0x06 chunk info.
// Chunks may be uncompressed (compression=none, what `rosbag record` writes by default), lz4 (`rosbag record --lz4`: an
// LZ4 frame, magic 0x184D2204) or bz2 (`rosbag record -j`: one bzip2 stream); both are decoded here by restatements of
// the published formats, because neither library may be assumed on the target -- checksums are not verified.
// tests/test_rosbag.py checks the decoders against chunks written by liblz4 and libbz2.
// Messages come back in record-time order (stable), like rosbag::View iterates them.
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "lego_loam_b200.h"

namespace {

struct Span {
  const uint8_t* p = nullptr;
  size_t n = 0;
};

struct Msg {
  uint64_t time_ns;
  size_t order;
  Span data;
};

thread_local std::string g_bag_error;

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// header fields -> name -> value bytes; false on a malformed header
bool parse_header(Span h, std::map<std::string, Span>* out) {
  size_t pos = 0;
  while (pos < h.n) {
    if (pos + 4 > h.n) return false;
    const uint32_t len = rd32(h.p + pos);
    pos += 4;
    if (len == 0 || pos + len > h.n) return false;
    const uint8_t* f = h.p + pos;
    const uint8_t* eq = (const uint8_t*)memchr(f, '=', len);
    if (!eq) return false;
    Span v;
    v.p = eq + 1;
    v.n = len - (size_t)(eq + 1 - f);
    (*out)[std::string((const char*)f, (size_t)(eq - f))] = v;
    pos += len;
  }
  return true;
}

bool next_record(Span buf, size_t* pos, Span* header, Span* data) {
  if (*pos + 4 > buf.n) return false;
  const uint32_t hl = rd32(buf.p + *pos);
  if (*pos + 4 + (size_t)hl + 4 > buf.n) return false;
  header->p = buf.p + *pos + 4;
  header->n = hl;
  const uint32_t dl = rd32(buf.p + *pos + 4 + hl);
  if (*pos + 8 + (size_t)hl + (size_t)dl > buf.n) return false;
  data->p = buf.p + *pos + 8 + hl;
  data->n = dl;
  *pos += 8 + (size_t)hl + (size_t)dl;
  return true;
}

std::string str_of(Span s) { return std::string((const char*)s.p, s.n); }

// ---- LZ4 (published "LZ4 Frame Format" 1.6 / "LZ4 Block Format") -----------------------------------------------------
// One block: sequences of <token><literal length ext><literals><u16 offset><match length ext>; the last sequence ends
// after its literals.  Output goes to dst[*op ...]; matches may reach back into earlier blocks (dependent blocks).
bool lz4_block(const uint8_t* ip, size_t n, uint8_t* dst, size_t cap, size_t* op) {
  const uint8_t* const iend = ip + n;
  while (ip < iend) {
    const unsigned token = *ip++;
    size_t lit = token >> 4;
    if (lit == 15) {
      unsigned b;
      do {
        if (ip >= iend) return false;
        b = *ip++;
        lit += b;
      } while (b == 255);
    }
    if ((size_t)(iend - ip) < lit || cap - *op < lit) return false;
    memcpy(dst + *op, ip, lit);
    ip += lit;
    *op += lit;
    if (ip >= iend) return true;  // last sequence: literals only
    if (iend - ip < 2) return false;
    const size_t offset = (size_t)ip[0] | ((size_t)ip[1] << 8);
    ip += 2;
    if (offset == 0 || offset > *op) return false;
    size_t mlen = token & 15u;
    if (mlen == 15) {
      unsigned b;
      do {
        if (ip >= iend) return false;
        b = *ip++;
        mlen += b;
      } while (b == 255);
    }
    mlen += 4;
    if (cap - *op < mlen) return false;
    for (size_t i = 0; i < mlen; ++i, ++*op) dst[*op] = dst[*op - offset];  // byte by byte: the match may overlap itself
  }
  return true;
}

// A whole frame into exactly `size` bytes (the chunk header's uncompressed size).
bool lz4_frame(const uint8_t* ip, size_t n, std::vector<uint8_t>* out, size_t size) {
  if (n < 7 || rd32(ip) != 0x184D2204u) return false;
  const unsigned flg = ip[4];
  if ((flg >> 6) != 1) return false;  // version 01
  const bool block_checksum = (flg >> 4) & 1, content_size = (flg >> 3) & 1, dict_id = flg & 1;
  size_t pos = 6 + (content_size ? 8 : 0) + (dict_id ? 4 : 0) + 1;  // FLG, BD, optional fields, header checksum
  out->resize(size);
  size_t op = 0;
  while (true) {
    if (pos + 4 > n) return false;
    const uint32_t bs = rd32(ip + pos);
    pos += 4;
    if (bs == 0) break;  // EndMark (a content checksum may follow)
    const size_t len = bs & 0x7fffffffu;
    if (pos + len > n) return false;
    if (bs & 0x80000000u) {  // stored block
      if (size - op < len) return false;
      memcpy(out->data() + op, ip + pos, len);
      op += len;
    } else if (!lz4_block(ip + pos, len, out->data(), size, &op)) {
      return false;
    }
    pos += len + (block_checksum ? 4 : 0);
  }
  return op == size;
}

// ---- bzip2 (restated from the published format description of bzip2 1.0.x: Huffman-coded MTF/RLE2 of a BWT block) ----
// Block and stream CRCs are not verified; the obsolete "randomised" block flag is refused.
struct BitReader {
  const uint8_t* p;
  size_t n, pos;  // pos in bits
  bool ok;
  uint32_t bits(int k) {  // MSB first, k <= 24
    uint32_t v = 0;
    for (int i = 0; i < k; ++i) {
      if ((pos >> 3) >= n) { ok = false; return 0; }
      v = (v << 1) | ((p[pos >> 3] >> (7 - (pos & 7))) & 1u);
      ++pos;
    }
    return v;
  }
};

bool bz2_stream(const uint8_t* ip, size_t n, std::vector<uint8_t>* out, size_t size) {
  if (n < 4 || ip[0] != 'B' || ip[1] != 'Z' || ip[2] != 'h' || ip[3] < '1' || ip[3] > '9') return false;
  const size_t block_max = (size_t)(ip[3] - '0') * 100000;
  BitReader br{ip, n, 32, true};
  out->clear();
  out->reserve(size);
  std::vector<uint32_t> tt(block_max);
  while (true) {
    const uint32_t m1 = br.bits(24), m2 = br.bits(24);
    if (!br.ok) return false;
    if (m1 == 0x177245u && m2 == 0x385090u) break;  // end of stream (combined CRC follows)
    if (m1 != 0x314159u || m2 != 0x265359u) return false;
    br.bits(16); br.bits(16);  // block CRC
    if (br.bits(1)) return false;  // randomised blocks: bzip2 < 0.9.5 only
    const uint32_t orig_ptr = br.bits(24);
    // symbols in use
    uint8_t seq_to_unseq[256];
    int n_in_use = 0;
    {
      const uint32_t ranges = br.bits(16);
      for (int i = 0; i < 16; ++i) {
        if (!((ranges >> (15 - i)) & 1u)) continue;
        const uint32_t used = br.bits(16);
        for (int j = 0; j < 16; ++j)
          if ((used >> (15 - j)) & 1u) seq_to_unseq[n_in_use++] = (uint8_t)(i * 16 + j);
      }
    }
    if (n_in_use == 0) return false;
    const int alpha = n_in_use + 2;
    const int n_groups = (int)br.bits(3);
    const int n_sel = (int)br.bits(15);
    if (n_groups < 2 || n_groups > 6 || n_sel < 1 || !br.ok) return false;
    std::vector<uint8_t> selector((size_t)n_sel);
    {
      uint8_t pos[6] = {0, 1, 2, 3, 4, 5};
      for (int i = 0; i < n_sel; ++i) {
        int j = 0;
        while (br.bits(1)) {
          if (++j >= n_groups) return false;
        }
        const uint8_t v = pos[j];
        for (; j > 0; --j) pos[j] = pos[j - 1];
        pos[0] = v;
        selector[(size_t)i] = v;
      }
    }
    // coding tables: delta-coded code lengths, then canonical Huffman decode tables
    int limit[6][24], base[6][24], perm[6][258], min_len[6];
    for (int t = 0; t < n_groups; ++t) {
      uint8_t len[258];
      int curr = (int)br.bits(5);
      for (int i = 0; i < alpha; ++i) {
        while (true) {
          if (curr < 1 || curr > 20 || !br.ok) return false;
          if (!br.bits(1)) break;
          curr += br.bits(1) ? -1 : 1;
        }
        len[i] = (uint8_t)curr;
      }
      int mn = 32, mx = 0;
      for (int i = 0; i < alpha; ++i) { mn = len[i] < mn ? len[i] : mn; mx = len[i] > mx ? len[i] : mx; }
      min_len[t] = mn;
      int pp = 0;
      for (int l = mn; l <= mx; ++l)
        for (int i = 0; i < alpha; ++i)
          if (len[i] == l) perm[t][pp++] = i;
      for (int i = 0; i < 24; ++i) { base[t][i] = 0; limit[t][i] = 0; }
      for (int i = 0; i < alpha; ++i) base[t][len[i] + 1]++;
      for (int i = 1; i < 24; ++i) base[t][i] += base[t][i - 1];
      int vec = 0;
      for (int l = mn; l <= mx; ++l) {
        vec += base[t][l + 1] - base[t][l];
        limit[t][l] = vec - 1;
        vec <<= 1;
      }
      for (int l = mn + 1; l <= mx; ++l) base[t][l] = ((limit[t][l - 1] + 1) << 1) - base[t][l];
      for (int l = mx + 1; l < 24; ++l) limit[t][l] = 0x7fffffff;  // a longer code cannot exist: stop the length walk
    }
    // MTF / RLE2 symbols -> tt
    const int eob = n_in_use + 1;
    uint8_t yy[256];
    for (int i = 0; i < 256; ++i) yy[i] = (uint8_t)i;
    uint32_t unzftab[256] = {0};
    size_t nblock = 0;
    int group_no = -1, group_pos = 0, t = 0;
    auto next_sym = [&]() -> int {
      if (group_pos == 0) {
        if (++group_no >= n_sel) { br.ok = false; return eob; }
        group_pos = 50;
        t = selector[(size_t)group_no];
      }
      --group_pos;
      int zn = min_len[t];
      int zvec = (int)br.bits(zn);
      while (zvec > limit[t][zn]) {
        if (++zn > 20 || !br.ok) { br.ok = false; return eob; }
        zvec = (zvec << 1) | (int)br.bits(1);
      }
      const int idx = zvec - base[t][zn];
      if (idx < 0 || idx >= alpha) { br.ok = false; return eob; }
      return perm[t][idx];
    };
    int sym = next_sym();
    while (br.ok && sym != eob) {
      if (sym <= 1) {  // RUNA / RUNB: a run length in bijective base 2
        size_t es = 0, weight = 1;
        do {
          es += weight << sym;  // RUNA adds weight, RUNB 2 * weight
          weight <<= 1;
          if (weight > (1u << 21)) return false;
          sym = next_sym();
        } while (br.ok && sym <= 1);
        const uint8_t uc = seq_to_unseq[yy[0]];
        if (nblock + es > block_max) return false;
        unzftab[uc] += (uint32_t)es;
        for (size_t i = 0; i < es; ++i) tt[nblock++] = uc;
        continue;
      }
      const int nn = sym - 1;
      const uint8_t v = yy[nn];
      for (int j = nn; j > 0; --j) yy[j] = yy[j - 1];
      yy[0] = v;
      const uint8_t uc = seq_to_unseq[v];
      if (nblock >= block_max) return false;
      unzftab[uc]++;
      tt[nblock++] = uc;
      sym = next_sym();
    }
    if (!br.ok || orig_ptr >= nblock) return false;
    // inverse BWT
    uint32_t cftab[257];
    cftab[0] = 0;
    for (int i = 0; i < 256; ++i) cftab[i + 1] = cftab[i] + unzftab[i];
    for (size_t i = 0; i < nblock; ++i) {
      const uint8_t uc = (uint8_t)(tt[i] & 0xffu);
      tt[cftab[uc]] |= (uint32_t)i << 8;
      cftab[uc]++;
    }
    uint32_t tpos = tt[orig_ptr] >> 8;
    // undo the initial run-length coding: four equal bytes are followed by a repeat count
    int run = 0, prev = -1;
    for (size_t i = 0; i < nblock; ++i) {
      tpos = tt[tpos];
      const uint8_t ch = (uint8_t)(tpos & 0xffu);
      tpos >>= 8;
      if (run == 4) {
        out->insert(out->end(), (size_t)ch, (uint8_t)prev);
        run = 0;
        prev = -1;
        continue;
      }
      out->push_back(ch);
      if ((int)ch == prev) {
        ++run;
      } else {
        run = 1;
        prev = ch;
      }
      if (out->size() > size) return false;
    }
  }
  return out->size() == size;
}

}  // namespace

struct ll_bag {
  std::vector<std::vector<uint8_t>> inflated;  // decompressed chunks that hold messages of the topic
  const uint8_t* file = nullptr;  // the whole bag, mapped read-only (recordings are gigabytes: no copy is made)
  size_t file_size = 0;
  std::string topic;
  std::map<uint32_t, std::pair<std::string, std::string>> conns;  // conn id -> (topic, type)
  std::vector<Msg> msgs;
};

namespace {

// records of one level (the file body or the data of a chunk)
int scan_records(ll_bag* bag, Span buf, bool top_level, const std::string& want_topic) {
  size_t pos = 0;
  while (pos < buf.n) {
    Span h, d;
    if (!next_record(buf, &pos, &h, &d)) { g_bag_error = "truncated record"; return LL_ERR_INVALID_ARG; }
    std::map<std::string, Span> f;
    if (!parse_header(h, &f) || !f.count("op") || f["op"].n != 1) { g_bag_error = "malformed record header"; return LL_ERR_INVALID_ARG; }
    const uint8_t op = f["op"].p[0];
    if (op == 0x05) {  // chunk
      if (!top_level) { g_bag_error = "chunk inside a chunk"; return LL_ERR_INVALID_ARG; }
      const std::string comp = f.count("compression") ? str_of(f["compression"]) : "none";
      if (comp == "lz4" || comp == "bz2") {
        if (!f.count("size") || f["size"].n != 4) { g_bag_error = "chunk record without size"; return LL_ERR_INVALID_ARG; }
        if (rd32(f["size"].p) > (1u << 30)) { g_bag_error = "chunk larger than 1 GiB"; return LL_ERR_INVALID_ARG; }  // rosbag's default is 768 kB
        bag->inflated.emplace_back();
        const bool good = comp == "lz4" ? lz4_frame(d.p, d.n, &bag->inflated.back(), rd32(f["size"].p))
                                        : bz2_stream(d.p, d.n, &bag->inflated.back(), rd32(f["size"].p));
        if (!good) { g_bag_error = "corrupt " + comp + " chunk"; return LL_ERR_INVALID_ARG; }
        Span u;
        u.p = bag->inflated.back().data();
        u.n = bag->inflated.back().size();
        const size_t before = bag->msgs.size();
        const int rc = scan_records(bag, u, false, want_topic);
        if (rc) return rc;
        if (bag->msgs.size() == before) bag->inflated.pop_back();  // nothing of the topic in this chunk: drop the copy
        continue;
      }
      if (comp != "none") { g_bag_error = "chunk compression '" + comp + "' is not supported (re-record or `rosbag decompress` the bag)"; return LL_ERR_INVALID_ARG; }
      const int rc = scan_records(bag, d, false, want_topic);
      if (rc) return rc;
    } else if (op == 0x07) {  // connection: header has conn + topic, data is a header with type, md5sum, message_definition
      if (!f.count("conn") || f["conn"].n != 4) { g_bag_error = "connection record without conn"; return LL_ERR_INVALID_ARG; }
      std::map<std::string, Span> cf;
      if (!parse_header(d, &cf)) { g_bag_error = "malformed connection data"; return LL_ERR_INVALID_ARG; }
      const std::string topic = f.count("topic") ? str_of(f["topic"]) : (cf.count("topic") ? str_of(cf["topic"]) : "");
      bag->conns[rd32(f["conn"].p)] = std::make_pair(topic, cf.count("type") ? str_of(cf["type"]) : "");
    } else if (op == 0x02) {  // message data
      if (!f.count("conn") || f["conn"].n != 4 || !f.count("time") || f["time"].n != 8) { g_bag_error = "message record without conn / time"; return LL_ERR_INVALID_ARG; }
      const auto it = bag->conns.find(rd32(f["conn"].p));
      if (it == bag->conns.end()) continue;  // the connection record always precedes its messages inside a chunk
      if (it->second.second != "sensor_msgs/PointCloud2") continue;
      if (!want_topic.empty() && it->second.first != want_topic) continue;
      if (want_topic.empty()) {
        if (bag->topic.empty()) bag->topic = it->second.first;  // first PointCloud2 topic of the bag
        if (it->second.first != bag->topic) continue;
      }
      Msg m;
      m.time_ns = (uint64_t)rd32(f["time"].p) * 1000000000ull + rd32(f["time"].p + 4);
      m.order = bag->msgs.size();
      m.data = d;
      bag->msgs.push_back(m);
    }
    // 0x03 bag header, 0x04 index data, 0x06 chunk info: not needed for a sequential read
  }
  return LL_OK;
}

struct Cursor {
  const uint8_t* p;
  size_t n, pos;
  bool ok;
  uint32_t u32() {
    if (pos + 4 > n) { ok = false; return 0; }
    const uint32_t v = rd32(p + pos);
    pos += 4;
    return v;
  }
  uint8_t u8() {
    if (pos + 1 > n) { ok = false; return 0; }
    return p[pos++];
  }
  Span bytes(size_t len) {
    Span s;
    if (pos + len > n) { ok = false; return s; }
    s.p = p + pos;
    s.n = len;
    pos += len;
    return s;
  }
};

}  // namespace

extern "C" {

const char* ll_bag_last_error(void) { return g_bag_error.c_str(); }

int ll_bag_open(const char* path, const char* topic, ll_bag** out) {
  if (!path || !out) return LL_ERR_INVALID_ARG;
  *out = nullptr;
  const int fd = open(path, O_RDONLY);
  struct stat sb;
  if (fd < 0 || fstat(fd, &sb) != 0) {  // main.cpp:31-34: ROS_FATAL + return 1
    if (fd >= 0) close(fd);
    g_bag_error = std::string("cannot open ") + path;
    return LL_ERR_INVALID_ARG;
  }
  ll_bag* bag = new ll_bag();
  bag->file_size = (size_t)sb.st_size;
  if (bag->file_size > 0) {
    void* m = mmap(nullptr, bag->file_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
      close(fd);
      delete bag;
      g_bag_error = std::string("cannot map ") + path;
      return LL_ERR_INVALID_ARG;
    }
    bag->file = (const uint8_t*)m;
  }
  close(fd);
  static const char magic[] = "#ROSBAG V2.0\n";
  if (bag->file_size < 13 || memcmp(bag->file, magic, 13) != 0) {
    g_bag_error = "not a rosbag v2.0 file";
    ll_bag_close(bag);
    return LL_ERR_INVALID_ARG;
  }
  bag->topic = topic ? topic : "";
  Span body;
  body.p = bag->file + 13;
  body.n = bag->file_size - 13;
  const int rc = scan_records(bag, body, true, bag->topic);
  if (rc) { ll_bag_close(bag); return rc; }
  std::stable_sort(bag->msgs.begin(), bag->msgs.end(), [](const Msg& a, const Msg& b) { return a.time_ns < b.time_ns; });
  *out = bag;
  return LL_OK;
}

int ll_bag_num_messages(const ll_bag* bag) { return bag ? (int)bag->msgs.size() : 0; }

const char* ll_bag_topic(const ll_bag* bag) { return bag ? bag->topic.c_str() : ""; }

// sensor_msgs/PointCloud2: Header header (uint32 seq, time stamp, string frame_id), uint32 height, uint32 width,
// PointField[] fields (string name, uint32 offset, uint8 datatype, uint32 count), bool is_bigendian, uint32 point_step,
// uint32 row_step, uint8[] data, bool is_dense; strings and arrays carry a uint32 length, everything little-endian
int ll_bag_get_pointcloud2(const ll_bag* bag, int index, ll_pointcloud2_view* v) {
  if (!bag || !v || index < 0 || index >= (int)bag->msgs.size()) return LL_ERR_INVALID_ARG;
  const Msg& m = bag->msgs[(size_t)index];
  Cursor c{m.data.p, m.data.n, 0, true};
  memset(v, 0, sizeof(*v));
  v->bag_time_ns = m.time_ns;
  c.u32();  // header.seq
  v->stamp_sec = c.u32();
  v->stamp_nsec = c.u32();
  c.bytes(c.u32());  // frame_id
  v->height = c.u32();
  v->width = c.u32();
  v->off_x = v->off_y = v->off_z = v->off_intensity = -1;
  const uint32_t nfields = c.u32();
  for (uint32_t i = 0; i < nfields && c.ok; ++i) {
    const Span name = c.bytes(c.u32());
    const uint32_t offset = c.u32();
    const uint8_t datatype = c.u8();
    c.u32();  // count
    if (!c.ok || datatype != 7) continue;  // FLOAT32
    const std::string nm = str_of(name);
    if (nm == "x") v->off_x = (int32_t)offset;
    else if (nm == "y") v->off_y = (int32_t)offset;
    else if (nm == "z") v->off_z = (int32_t)offset;
    else if (nm == "intensity") v->off_intensity = (int32_t)offset;
  }
  v->is_bigendian = c.u8();
  v->point_step = c.u32();
  v->row_step = c.u32();
  const Span data = c.bytes(c.u32());
  v->is_dense = c.u8();
  if (!c.ok) { g_bag_error = "truncated sensor_msgs/PointCloud2 message"; return LL_ERR_INVALID_ARG; }
  v->data = data.p;
  v->data_len = data.n;
  return LL_OK;
}

void ll_bag_close(ll_bag* bag) {
  if (!bag) return;
  if (bag->file) munmap((void*)bag->file, bag->file_size);
  delete bag;
}

}  // extern "C"
