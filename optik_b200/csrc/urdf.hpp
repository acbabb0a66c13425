// urdf.hpp -- minimal XML reader + URDF -> serial kinematic chain, host side.
//
// Restates the model-loading semantics of the reference (kylc/optik @ 355e463):
//   parse_urdf               crates/optik/src/kinematics.rs:269-319
//   urdf_to_tfm (xyz + rpy)  crates/optik/src/kinematics.rs:263-267
//   KinematicChain::from_urdf crates/optik/src/kinematics.rs:18-105
// (links = graph nodes, joints = directed parent->child edges, acyclicity check, shortest base->EE path,
//  fixed joints folded into the next articulated joint with the reference's `joint.origin * collapsed`
//  product order, trailing fixed joints become one fixed tip joint, empty chains rejected.)
// Attribute defaults follow urdf-rs 0.9: origin xyz/rpy = 0, axis = (1,0,0), limit lower = upper = 0.
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace optik {

// ------------------------------------------------------------------ tiny XML
struct XmlNode {
  std::string name;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<XmlNode>> children;
  const XmlNode* child(const char* n) const {
    for (auto& c : children)
      if (c->name == n) return c.get();
    return nullptr;
  }
};

class XmlReader {
 public:
  explicit XmlReader(const std::string& s) : s_(s) {}
  std::unique_ptr<XmlNode> parse() {
    skip_misc();
    auto root = element();
    if (!root) fail("no root element");
    return root;
  }

 private:
  const std::string& s_;
  size_t p_ = 0;
  [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("error parsing URDF file! (") + what + ")"); }
  bool starts(const char* lit) const { return s_.compare(p_, strlen(lit), lit) == 0; }
  void skip_ws() {
    while (p_ < s_.size() && isspace((unsigned char)s_[p_])) p_++;
  }
  void skip_until(const char* end) {
    size_t e = s_.find(end, p_);
    if (e == std::string::npos) fail("unterminated markup");
    p_ = e + strlen(end);
  }
  void skip_misc() {  // whitespace, <?...?>, <!--...-->, <!DOCTYPE ...>
    for (;;) {
      skip_ws();
      if (starts("<?")) skip_until("?>");
      else if (starts("<!--")) skip_until("-->");
      else if (starts("<!")) skip_until(">");
      else return;
    }
  }
  static bool name_char(char c) { return isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }
  std::string name() {
    size_t b = p_;
    while (p_ < s_.size() && name_char(s_[p_])) p_++;
    if (b == p_) fail("expected a name");
    return s_.substr(b, p_ - b);
  }
  static std::string unescape(const std::string& v) {
    std::string o;
    for (size_t i = 0; i < v.size(); i++) {
      if (v[i] != '&') { o += v[i]; continue; }
      static const char* ent[][2] = {{"&quot;", "\""}, {"&apos;", "'"}, {"&lt;", "<"}, {"&gt;", ">"}, {"&amp;", "&"}};
      bool hit = false;
      for (auto& e : ent)
        if (v.compare(i, strlen(e[0]), e[0]) == 0) { o += e[1]; i += strlen(e[0]) - 1; hit = true; break; }
      if (!hit) o += v[i];
    }
    return o;
  }
  std::unique_ptr<XmlNode> element() {
    if (p_ >= s_.size() || s_[p_] != '<') return nullptr;
    p_++;
    auto node = std::make_unique<XmlNode>();
    node->name = name();
    for (;;) {
      skip_ws();
      if (p_ >= s_.size()) fail("unterminated tag");
      if (starts("/>")) { p_ += 2; return node; }
      if (s_[p_] == '>') { p_++; break; }
      std::string key = name();
      skip_ws();
      if (p_ >= s_.size() || s_[p_] != '=') fail("expected '='");
      p_++;
      skip_ws();
      if (p_ >= s_.size() || (s_[p_] != '"' && s_[p_] != '\'')) fail("expected a quoted value");
      char qc = s_[p_++];
      size_t e = s_.find(qc, p_);
      if (e == std::string::npos) fail("unterminated attribute");
      node->attr[key] = unescape(s_.substr(p_, e - p_));
      p_ = e + 1;
    }
    for (;;) {  // content
      size_t lt = s_.find('<', p_);
      if (lt == std::string::npos) fail("unterminated element");
      p_ = lt;
      if (starts("<!--")) { skip_until("-->"); continue; }
      if (starts("<![CDATA[")) { skip_until("]]>"); continue; }
      if (starts("<?")) { skip_until("?>"); continue; }
      if (starts("</")) {
        p_ += 2;
        std::string close = name();
        if (close != node->name) fail("mismatched closing tag");
        skip_ws();
        if (p_ >= s_.size() || s_[p_] != '>') fail("malformed closing tag");
        p_++;
        return node;
      }
      node->children.push_back(element());
    }
  }
};

// ------------------------------------------------------------------ chain
enum JointType { REVOLUTE = 0, PRISMATIC = 1, FIXED = 2 };

struct Pose {  // Isometry3<f64>
  double q[4] = {0, 0, 0, 1};  // xyzw
  double t[3] = {0, 0, 0};
  bool is_identity() const { return q[0] == 0 && q[1] == 0 && q[2] == 0 && q[3] == 1 && t[0] == 0 && t[1] == 0 && t[2] == 0; }
};
inline Pose pose_mul(const Pose& a, const Pose& b) {
  Pose o;
  const double ax = a.q[0], ay = a.q[1], az = a.q[2], aw = a.q[3];
  const double bx = b.q[0], by = b.q[1], bz = b.q[2], bw = b.q[3];
  o.q[0] = aw * bx + ax * bw + ay * bz - az * by;
  o.q[1] = aw * by - ax * bz + ay * bw + az * bx;
  o.q[2] = aw * bz + ax * by - ay * bx + az * bw;
  o.q[3] = aw * bw - ax * bx - ay * by - az * bz;
  // rotate b.t by a.q:  v + 2w(u x v) + 2u x (u x v)
  const double vx = b.t[0], vy = b.t[1], vz = b.t[2];
  const double cx = ay * vz - az * vy, cy = az * vx - ax * vz, cz = ax * vy - ay * vx;
  const double dx = ay * cz - az * cy, dy = az * cx - ax * cz, dz = ax * cy - ay * cx;
  o.t[0] = a.t[0] + vx + 2 * (aw * cx + dx);
  o.t[1] = a.t[1] + vy + 2 * (aw * cy + dy);
  o.t[2] = a.t[2] + vz + 2 * (aw * cz + dz);
  return o;
}

struct Joint {
  std::string name;
  int type = FIXED;
  double axis[3] = {0, 0, 0};
  double lower = 0, upper = 0;  // only for articulated joints; may be -inf/+inf
  Pose origin;
};

namespace detail {
inline void parse_vec(const std::string& s, double* out, int n) {
  const char* c = s.c_str();
  for (int i = 0; i < n; i++) {
    char* end = nullptr;
    out[i] = strtod(c, &end);
    if (end == c) throw std::runtime_error("error parsing URDF file! (bad vector '" + s + "')");
    c = end;
  }
}
inline Pose pose_from_xyz_rpy(const double* xyz, const double* rpy) {  // kinematics.rs:263-267
  Pose p;
  const double sr = sin(rpy[0] / 2), cr = cos(rpy[0] / 2), sp = sin(rpy[1] / 2), cp = cos(rpy[1] / 2),
               sy = sin(rpy[2] / 2), cy = cos(rpy[2] / 2);
  p.q[0] = sr * cp * cy - cr * sp * sy;
  p.q[1] = cr * sp * cy + sr * cp * sy;
  p.q[2] = cr * cp * sy - sr * sp * cy;
  p.q[3] = cr * cp * cy + sr * sp * sy;
  p.t[0] = xyz[0]; p.t[1] = xyz[1]; p.t[2] = xyz[2];
  return p;
}
struct Edge {
  int parent, child;
  Joint joint;
};
}  // namespace detail

// Throws std::runtime_error with the reference's panic messages.
inline std::vector<Joint> chain_from_urdf(const std::string& text, const std::string& base_link,
                                          const std::string& ee_link, bool urdf_correct_fold) {
  auto root = XmlReader(text).parse();
  if (root->name != "robot") throw std::runtime_error("error parsing URDF file! (root element is not <robot>)");
  std::vector<std::string> links;
  for (auto& c : root->children)
    if (c->name == "link") {
      auto it = c->attr.find("name");
      if (it == c->attr.end()) throw std::runtime_error("error parsing URDF file! (link without a name)");
      links.push_back(it->second);
    }
  auto link_ix = [&](const std::string& n) {
    for (size_t i = 0; i < links.size(); i++)
      if (links[i] == n) return (int)i;
    return -1;
  };
  std::vector<detail::Edge> edges;
  for (auto& c : root->children) {
    if (c->name != "joint") continue;
    const XmlNode* parent = c->child("parent");
    const XmlNode* child = c->child("child");
    if (!parent || !child || !parent->attr.count("link") || !child->attr.count("link") || !c->attr.count("type"))
      throw std::runtime_error("error parsing URDF file! (joint without parent/child/type)");
    detail::Edge e;
    e.parent = link_ix(parent->attr.at("link"));
    if (e.parent < 0) throw std::runtime_error("joint parent link '" + parent->attr.at("link") + "' does not exist");
    e.child = link_ix(child->attr.at("link"));
    if (e.child < 0) throw std::runtime_error("joint child link '" + child->attr.at("link") + "' does not exist");
    const std::string& typ = c->attr.at("type");
    if (typ == "revolute") e.joint.type = REVOLUTE;
    else if (typ == "prismatic") e.joint.type = PRISMATIC;
    else if (typ == "fixed") e.joint.type = FIXED;
    else throw std::runtime_error("joint type not supported: " + typ);
    e.joint.name = c->attr.count("name") ? c->attr.at("name") : "";
    double xyz[3] = {0, 0, 0}, rpy[3] = {0, 0, 0}, axis[3] = {1, 0, 0};
    if (const XmlNode* o = c->child("origin")) {
      if (o->attr.count("xyz")) detail::parse_vec(o->attr.at("xyz"), xyz, 3);
      if (o->attr.count("rpy")) detail::parse_vec(o->attr.at("rpy"), rpy, 3);
    }
    if (const XmlNode* a = c->child("axis"))
      if (a->attr.count("xyz")) detail::parse_vec(a->attr.at("xyz"), axis, 3);
    double lo = 0, hi = 0;
    if (const XmlNode* l = c->child("limit")) {
      if (l->attr.count("lower")) detail::parse_vec(l->attr.at("lower"), &lo, 1);
      if (l->attr.count("upper")) detail::parse_vec(l->attr.at("upper"), &hi, 1);
    }
    e.joint.origin = detail::pose_from_xyz_rpy(xyz, rpy);
    if (e.joint.type != FIXED) {
      const double nrm = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
      for (int i = 0; i < 3; i++) e.joint.axis[i] = axis[i] / nrm;
      if (hi - lo > 0.0) { e.joint.lower = lo; e.joint.upper = hi; }           // kinematics.rs:299-303
      else { e.joint.lower = -INFINITY; e.joint.upper = INFINITY; }
    }
    edges.push_back(e);
  }
  const int L = (int)links.size();
  std::vector<std::vector<int>> out(L);
  for (size_t i = 0; i < edges.size(); i++) out[edges[i].parent].push_back((int)i);
  // acyclicity (kinematics.rs:21): iterative three-colour DFS
  {
    std::vector<int> colour(L, 0), it(L, 0), stack;
    for (int s = 0; s < L; s++) {
      if (colour[s]) continue;
      stack.push_back(s);
      colour[s] = 1;
      while (!stack.empty()) {
        int u = stack.back();
        if (it[u] < (int)out[u].size()) {
          int v = edges[out[u][it[u]++]].child;
          if (colour[v] == 1) throw std::runtime_error("robot model contains loops");
          if (colour[v] == 0) { colour[v] = 1; stack.push_back(v); }
        } else {
          colour[u] = 2;
          stack.pop_back();
        }
      }
    }
  }
  const int b = link_ix(base_link), e = link_ix(ee_link);
  if (b < 0) throw std::runtime_error("base link '" + base_link + "' does not exist");
  if (e < 0) throw std::runtime_error("EE link '" + ee_link + "' does not exist");
  // shortest path by hop count (A* with unit edge cost and zero heuristic, kinematics.rs:35-42)
  std::vector<int> via(L, -1), seen(L, 0), frontier{b};
  seen[b] = 1;
  while (!frontier.empty() && !seen[e]) {
    std::vector<int> next;
    for (int u : frontier)
      for (int ei : out[u]) {
        int v = edges[ei].child;
        if (!seen[v]) { seen[v] = 1; via[v] = ei; next.push_back(v); }
      }
    frontier.swap(next);
  }
  if (!seen[e]) throw std::runtime_error("no path from base to EE link");
  std::vector<int> path;
  for (int l = e; l != b; l = edges[via[l]].parent) path.push_back(via[l]);
  // fold fixed joints (kinematics.rs:64-86)
  std::vector<Joint> chain;
  Pose collapsed;
  for (auto it = path.rbegin(); it != path.rend(); ++it) {
    const Joint& j = edges[*it].joint;
    Pose folded = urdf_correct_fold ? pose_mul(collapsed, j.origin) : pose_mul(j.origin, collapsed);
    if (j.type == FIXED) {
      collapsed = folded;
    } else {
      Joint nj = j;
      nj.origin = folded;
      chain.push_back(nj);
      collapsed = Pose();
    }
  }
  if (!collapsed.is_identity()) {  // kinematics.rs:90-97
    Joint tip;
    tip.type = FIXED;
    tip.origin = collapsed;
    chain.push_back(tip);
  }
  int nq = 0;
  for (auto& j : chain) nq += (j.type != FIXED);
  if (nq == 0) throw std::runtime_error("kinematic chain is empty");  // kinematics.rs:102
  return chain;
}

}  // namespace optik
