// gl_harness.c -- TEST INFRASTRUCTURE.  Runs the reference's own visualisation pass -- its UNMODIFIED shader text
// (shader/voxel_cone_tracing.vert / .frag, read from the reference tree where it lies) and the GL state of
// Renderer::visualize / draw_models / upload_lights / upload_camera / create_tex_3d (src/renderer.cpp:240-281,355-390,
// src/texture_3d.cpp:3-25) -- on a real OpenGL implementation: the Mesa 18.1 llvmpipe that ships inside Nsight Compute
// (the only GL in this image).  That library is a GLX/xlib build; oracle/gl_ref/fake_x11.c stands in for libX11 so that it
// creates an off-screen context without an X server, and MESA_GL(SL)_VERSION_OVERRIDE make it accept "#version 450".
// llvmpipe 18.1 has geometry shaders but no image load/store and no compute shaders, so the voxelization and mip passes
// cannot run on it: the six voxel textures are filled from a file (the oracle's grid + mip chain) and what this harness pins
// is everything GL does *around and in* the cone tracer: rasterisation, clipping, interpolation, depth test, textureLod
// filtering, blending, the unorm conversion, and the fragment shader itself.
//
//   vct_gl_ref <shader dir> <job file> <out file>
// job file (little endian; written by oracle/gl_ref.py):
//   "VCTGLJOB" u32 W H R levels n_verts n_indices n_draws n_mats n_lights  i32 direct diffuse specular shadow view_voxel_dir
//   f32 view_voxel_lod cube_size  f32 view[16] proj[16]  lights n*7 f32  materials n*128 B  draws n*80 B (first, count, vertex_base,
//   material, model[16])  vertices n*32 B  indices n*4 B  level 0: R^3 u32  then for dir 0..5, level 1..levels-1: (R>>l)^3 u32
// out file: RGBA8 frame W*H*4 (row 0 = bottom), then the same frame rendered into an RGBA32F target W*H*16
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "fake_x11.h"

typedef void* GLXFBConfig;
typedef void* GLXContext;
typedef XID GLXPbuffer;
typedef unsigned GLenum, GLuint, GLbitfield;
typedef int GLint, GLsizei;
typedef float GLfloat;
typedef ptrdiff_t GLsizeiptr, GLintptr;

#define GL_FUNCS(X)                                                                                                          \
  X(GLuint, CreateShader, (GLenum)) X(void, ShaderSource, (GLuint, GLsizei, const char**, const GLint*)) X(void, CompileShader, (GLuint))       \
  X(void, GetShaderiv, (GLuint, GLenum, GLint*)) X(void, GetShaderInfoLog, (GLuint, GLsizei, GLsizei*, char*)) X(GLuint, CreateProgram, (void))  \
  X(void, AttachShader, (GLuint, GLuint)) X(void, LinkProgram, (GLuint)) X(void, GetProgramiv, (GLuint, GLenum, GLint*))                        \
  X(void, GetProgramInfoLog, (GLuint, GLsizei, GLsizei*, char*)) X(void, UseProgram, (GLuint)) X(GLint, GetUniformLocation, (GLuint, const char*)) \
  X(GLuint, GetUniformBlockIndex, (GLuint, const char*)) X(void, Uniform1i, (GLint, GLint)) X(void, Uniform1f, (GLint, GLfloat))                \
  X(void, Uniform3fv, (GLint, GLsizei, const GLfloat*)) X(void, UniformMatrix4fv, (GLint, GLsizei, unsigned char, const GLfloat*))              \
  X(void, GenBuffers, (GLsizei, GLuint*)) X(void, BindBuffer, (GLenum, GLuint)) X(void, BufferData, (GLenum, GLsizeiptr, const void*, GLenum))   \
  X(void, BindBufferBase, (GLenum, GLuint, GLuint)) X(void, GenVertexArrays, (GLsizei, GLuint*)) X(void, BindVertexArray, (GLuint))             \
  X(void, EnableVertexAttribArray, (GLuint)) X(void, VertexAttribPointer, (GLuint, GLint, GLenum, unsigned char, GLsizei, const void*))         \
  X(void, GenTextures, (GLsizei, GLuint*)) X(void, BindTexture, (GLenum, GLuint)) X(void, TexParameteri, (GLenum, GLenum, GLint))               \
  X(void, TexStorage3D, (GLenum, GLsizei, GLenum, GLsizei, GLsizei, GLsizei))                                                                   \
  X(void, TexSubImage3D, (GLenum, GLint, GLint, GLint, GLint, GLsizei, GLsizei, GLsizei, GLenum, GLenum, const void*))                          \
  X(void, ActiveTexture, (GLenum)) X(void, GenFramebuffers, (GLsizei, GLuint*)) X(void, BindFramebuffer, (GLenum, GLuint))                      \
  X(void, GenRenderbuffers, (GLsizei, GLuint*)) X(void, BindRenderbuffer, (GLenum, GLuint))                                                     \
  X(void, RenderbufferStorage, (GLenum, GLenum, GLsizei, GLsizei)) X(void, FramebufferRenderbuffer, (GLenum, GLenum, GLenum, GLuint))           \
  X(GLenum, CheckFramebufferStatus, (GLenum)) X(void, Viewport, (GLint, GLint, GLsizei, GLsizei)) X(void, Enable, (GLenum))                     \
  X(void, Disable, (GLenum)) X(void, Clear, (GLbitfield)) X(void, ClearColor, (GLfloat, GLfloat, GLfloat, GLfloat))                             \
  X(void, BlendFunc, (GLenum, GLenum)) X(void, DrawElementsBaseVertex, (GLenum, GLsizei, GLenum, const void*, GLint))                           \
  X(void, ReadPixels, (GLint, GLint, GLsizei, GLsizei, GLenum, GLenum, void*)) X(void, PixelStorei, (GLenum, GLint)) X(GLenum, GetError, (void)) \
  X(void, DrawBuffers, (GLsizei, const GLenum*)) X(void, ReadBuffer, (GLenum)) X(void, ColorMask, (unsigned char, unsigned char, unsigned char, unsigned char)) \
  X(void, FramebufferTextureLayer, (GLenum, GLenum, GLuint, GLint, GLint)) X(void, GetTexImage, (GLenum, GLint, GLenum, GLenum, void*)) \
  X(void, DrawArrays, (GLenum, GLint, GLsizei)) X(void, Finish, (void)) X(const unsigned char*, GetString, (GLenum)) X(void, ClampColor, (GLenum, GLenum))

#define X(ret, name, args) static ret(*gl##name) args;
GL_FUNCS(X)
#undef X

static void die(const char* what) { fprintf(stderr, "vct_gl_ref: %s\n", what); exit(1); }

static char* slurp(const char* dir, const char* name) {
  char path[4096];
  snprintf(path, sizeof path, "%s/%s", dir, name);
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "vct_gl_ref: cannot read %s\n", path); exit(1); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  rewind(f);
  char* s = malloc(n + 1);
  if (fread(s, 1, n, f) != (size_t)n) die("short read");
  s[n] = 0;
  fclose(f);
  return s;
}

static GLuint compile(GLenum type, const char* src, const char* name) {
  GLuint sh = glCreateShader(type);
  glShaderSource(sh, 1, &src, 0);
  glCompileShader(sh);
  GLint ok = 0;
  glGetShaderiv(sh, 0x8B81 /*COMPILE_STATUS*/, &ok);
  if (!ok) {
    char log[8192] = "";
    glGetShaderInfoLog(sh, sizeof log, 0, log);
    fprintf(stderr, "vct_gl_ref: %s does not compile:\n%s\n", name, log);
    exit(1);
  }
  return sh;
}

struct Job {
  uint32_t W, H, R, levels, n_verts, n_indices, n_draws, n_mats, n_lights;
  int32_t direct, diffuse, specular, shadow, view_voxel_dir;
  float view_voxel_lod, cube_size, view[16], proj[16];
};

// ---- Renderer::voxelize (renderer.cpp:316-353) as far as this driver can run it: the reference's voxelize.vert and voxelize.geom
// unmodified, and its voxelize.frag with the image store replaced by two colour outputs (oracle/gl_ref.py does that to the text: llvmpipe
// 18.1 has no image load/store) -- voxel coordinate and colour of every fragment, i.e. the arguments of imageAtomicRGBA8Avg.  GL state as
// the reference sets it: viewport 2R x 2R, no depth test, no culling, no blending.  One draw per triangle, so that the fragments come
// out as a list in draw order / triangle order / row / column.
// out file: u32 n, then n records of 10 words: u32 triangle sequence, u32 x, u32 y, f32 voxel[3], f32 colour[4]
static int voxelize_pass(const struct Job* J, char** argv, const float* lights, const unsigned char* draws, const GLuint* mat_ubo, GLuint cam_ubo) {
  GLuint prog = glCreateProgram();
  glAttachShader(prog, compile(0x8B31, slurp(argv[1], "voxelize.vert"), "voxelize.vert"));
  glAttachShader(prog, compile(0x8DD9, slurp(argv[1], "voxelize.geom"), "voxelize.geom"));
  glAttachShader(prog, compile(0x8B30, slurp(argv[1], "voxelize.frag"), "voxelize.frag (image store replaced by colour outputs)"));
  glLinkProgram(prog);
  GLint ok = 0;
  glGetProgramiv(prog, 0x8B82, &ok);
  if (!ok) { char log[8192] = ""; glGetProgramInfoLog(prog, sizeof log, 0, log); fprintf(stderr, "%s\n", log); die("link failed (voxelize)"); }
  const GLuint material_location = glGetUniformBlockIndex(prog, "material"), camera_location = glGetUniformBlockIndex(prog, "camera");
  fprintf(stderr, "voxelize program: uniform block indices camera %u, material %u\n", camera_location, material_location);
  const GLsizei V = (GLsizei)(2 * J->R);   // viewport_res = m_resolution * 2
  GLuint fbo, rb[2];
  glGenFramebuffers(1, &fbo); glBindFramebuffer(0x8D40, fbo);
  glGenRenderbuffers(2, rb);
  for (int i = 0; i < 2; i++) {
    glBindRenderbuffer(0x8D41, rb[i]); glRenderbufferStorage(0x8D41, 0x8814 /*RGBA32F*/, V, V);
    glFramebufferRenderbuffer(0x8D40, 0x8CE0 + i, 0x8D41, rb[i]);
  }
  const GLenum bufs[2] = {0x8CE0, 0x8CE1};
  glDrawBuffers(2, bufs);
  if (glCheckFramebufferStatus(0x8D40) != 0x8CD5) die("framebuffer incomplete (voxelize)");
  glClampColor(0x891B, 0); glClampColor(0x891C, 0);
  glUseProgram(prog);
  glUniform1f(glGetUniformLocation(prog, "cube_size"), J->cube_size);
  glUniform1i(glGetUniformLocation(prog, "vct_grid_res"), (GLint)J->R);
  glBindBufferBase(0x8A11, camera_location, cam_ubo);
  glUniform3fv(glGetUniformLocation(prog, "camera_position"), 1, J->view + 12);
  for (uint32_t i = 0; i < J->n_lights; i++) {
    char name[64];
    snprintf(name, sizeof name, "point_lights[%u].position", i); glUniform3fv(glGetUniformLocation(prog, name), 1, lights + 7 * i);
    snprintf(name, sizeof name, "point_lights[%u].color", i); glUniform3fv(glGetUniformLocation(prog, name), 1, lights + 7 * i + 3);
    snprintf(name, sizeof name, "point_lights[%u].intensity", i); glUniform1f(glGetUniformLocation(prog, name), lights[7 * i + 6]);
  }
  glUniform1i(glGetUniformLocation(prog, "point_light_count"), (GLint)J->n_lights);
  glViewport(0, 0, V, V);
  glDisable(0x0B44 /*CULL_FACE*/);
  glDisable(0x0B71 /*DEPTH_TEST*/);
  glDisable(0x0BE2 /*BLEND*/);
  glClearColor(0, 0, 0, 0);
  glPixelStorei(0x0D05, 1);
  FILE* out = fopen(argv[3], "wb");
  if (!out) die("cannot open the output file");
  uint32_t n_frag = 0, seq = 0;
  fwrite(&n_frag, 4, 1, out);
  float* vox = malloc((size_t)V * V * 16);
  float* col = malloc((size_t)V * V * 16);
  const GLint model_location = glGetUniformLocation(prog, "model");
  for (uint32_t i = 0; i < J->n_draws; i++) {
    uint32_t d[4];
    memcpy(d, draws + 80 * (size_t)i, 16);
    glUniformMatrix4fv(model_location, 1, 0, (const float*)(draws + 80 * (size_t)i + 16));
    glBindBufferBase(0x8A11, material_location, mat_ubo[d[3]]);
    for (uint32_t t = 0; t + 3 <= d[1]; t += 3, seq++) {
      glClear(0x4000);
      glDrawElementsBaseVertex(0x0004, 3, 0x1405, (const void*)(sizeof(unsigned) * ((size_t)d[0] + t)), (GLint)d[2]);
      glReadBuffer(0x8CE0); glReadPixels(0, 0, V, V, 0x1908, 0x1406, vox);
      glReadBuffer(0x8CE1); glReadPixels(0, 0, V, V, 0x1908, 0x1406, col);
      for (GLsizei y = 0; y < V; y++)
        for (GLsizei x = 0; x < V; x++) {
          const float* v = vox + 4 * ((size_t)y * V + x);
          if (v[3] != 1.0f) continue;
          const uint32_t head[3] = {seq, (uint32_t)x, (uint32_t)y};
          fwrite(head, 4, 3, out);
          fwrite(v, 4, 3, out);
          fwrite(col + 4 * ((size_t)y * V + x), 4, 4, out);
          n_frag++;
        }
    }
  }
  if (glGetError()) die("GL error in the voxelization pass");
  fseek(out, 0, SEEK_SET);
  fwrite(&n_frag, 4, 1, out);
  fclose(out);
  fprintf(stderr, "voxelize: %u triangles, %u fragments\n", seq, n_frag);
  return 0;
}

// ---- Renderer::filter (renderer.cpp:283-314) as far as this driver can run it: llvmpipe 18.1 has no compute shaders, so the reference's
// mipmap.comp runs as a FRAGMENT shader (oracle/gl_ref.py rewrites three lines of its text: the invocation id comes from gl_FragCoord and a
// layer uniform, the six imageStores become six colour outputs); its arithmetic, its texelFetches and its constant tables are untouched.
// One draw per z layer of the destination level into six RGBA8 attachments (layer z of level mip + 1 of the six textures), so the
// float -> unorm8 conversion is GL's own, like the image store's.
// mip job: "VCTGLMIP" u32 R levels, R^3 u32 (level 0).  out file: for dir 0..5, level 1..levels-1: (R >> l)^3 u32
static int mip_pass(char** argv) {
  FILE* f = fopen(argv[2], "rb");
  if (!f) die("cannot open the job file");
  char magic[8];
  uint32_t hdr[2];
  if (fread(magic, 1, 8, f) != 8 || fread(hdr, 4, 2, f) != 2) die("bad mip job");
  const uint32_t R = hdr[0], levels = hdr[1];
  const size_t n0 = (size_t)R * R * R;
  unsigned char* level0 = malloc(n0 * 4);
  if (fread(level0, 4, n0, f) != n0) die("short mip job");
  fclose(f);
  GLuint tex[6];
  glGenTextures(6, tex);
  glPixelStorei(0x0CF5, 1); glPixelStorei(0x0D05, 1);
  for (int d = 0; d < 6; d++) {   // create_tex_3d
    glActiveTexture(0x84C0 + d);   // activate_tex_3d(m_mipmap_shader, m_voxel_maps[i], i)
    glBindTexture(0x806F, tex[d]);
    glTexParameteri(0x806F, 0x2802, 0x812D); glTexParameteri(0x806F, 0x2803, 0x812D); glTexParameteri(0x806F, 0x8072, 0x812D);
    glTexParameteri(0x806F, 0x2801, 0x2703);
    glTexStorage3D(0x806F, (GLsizei)levels, 0x8058, R, R, R);
    glTexSubImage3D(0x806F, 0, 0, 0, 0, R, R, R, 0x1908, 0x1401, level0);
  }
  if (glGetError()) die("GL error while creating the textures (mip)");
  static const char* vs =
      "#version 450 core\nvoid main() { vec2 p = vec2((gl_VertexID << 1) & 2, gl_VertexID & 2); gl_Position = vec4(p * 2.0 - 1.0, 0.0, 1.0); }\n";
  GLuint prog = glCreateProgram();
  glAttachShader(prog, compile(0x8B31, vs, "full-screen triangle"));
  glAttachShader(prog, compile(0x8B30, slurp(argv[1], "mipmap.comp"), "mipmap.comp (as a fragment shader)"));
  glLinkProgram(prog);
  GLint ok = 0;
  glGetProgramiv(prog, 0x8B82, &ok);
  if (!ok) { char log[8192] = ""; glGetProgramInfoLog(prog, sizeof log, 0, log); fprintf(stderr, "%s\n", log); die("link failed (mip)"); }
  glUseProgram(prog);
  GLuint vao, fbo;
  glGenVertexArrays(1, &vao); glBindVertexArray(vao);
  glGenFramebuffers(1, &fbo); glBindFramebuffer(0x8D40, fbo);
  const GLenum bufs[6] = {0x8CE0, 0x8CE1, 0x8CE2, 0x8CE3, 0x8CE4, 0x8CE5};
  glDisable(0x0B71); glDisable(0x0BE2); glDisable(0x0B44);
  const GLint resolution_location = glGetUniformLocation(prog, "resolution"), mip_location = glGetUniformLocation(prog, "mip"),
              layer_location = glGetUniformLocation(prog, "vct_layer");
  uint32_t current_dim = R;
  for (uint32_t mip = 0; mip + 1 < levels; mip++, current_dim /= 2) {   // filter(): while (current_dim >= 1), levels the texture has
    const GLsizei N = (GLsizei)(current_dim / 2);
    glUniform1i(resolution_location, (GLint)current_dim);
    glUniform1i(mip_location, (GLint)mip);
    glViewport(0, 0, N, N);
    for (GLsizei z = 0; z < N; z++) {
      for (int d = 0; d < 6; d++) glFramebufferTextureLayer(0x8D40, 0x8CE0 + d, tex[d], (GLint)mip + 1, z);
      glDrawBuffers(6, bufs);
      if (glCheckFramebufferStatus(0x8D40) != 0x8CD5) die("framebuffer incomplete (mip)");
      glUniform1i(layer_location, z);
      glDrawArrays(0x0004, 0, 3);
    }
    glFinish();
  }
  if (glGetError()) die("GL error in the mip pass");
  FILE* out = fopen(argv[3], "wb");
  if (!out) die("cannot open the output file");
  for (int d = 0; d < 6; d++) {
    glActiveTexture(0x84C0 + d);
    for (uint32_t l = 1; l < levels; l++) {
      const size_t N = R >> l;
      void* buf = malloc(N * N * N * 4 + 16);
      glGetTexImage(0x806F, (GLint)l, 0x1908, 0x1401, buf);
      fwrite(buf, 4, N * N * N, out);
      free(buf);
    }
  }
  if (glGetError()) die("GL error while reading the levels back");
  fclose(out);
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 4) die("usage: vct_gl_ref <shader dir> <job file> <out file>");
  const char* libgl = getenv("VCT_MESA_LIBGL");
  if (!libgl) libgl = "/opt/nvidia/nsight-compute/2025.2.1/host/linux-desktop-glibc_2_11_3-x64/Mesa/libGL.so.1";
  setenv("MESA_GL_VERSION_OVERRIDE", "4.5", 1);
  setenv("MESA_GLSL_VERSION_OVERRIDE", "450", 1);
  // the stand-in libX11 / libXext sit next to this binary; loaded first, their SONAMEs satisfy libGL's DT_NEEDED entries
  char self[4096], path[4200];
  ssize_t n = readlink("/proc/self/exe", self, sizeof self - 1);
  if (n <= 0) die("readlink /proc/self/exe");
  self[n] = 0;
  *strrchr(self, '/') = 0;
  snprintf(path, sizeof path, "%s/libX11.so.6", self);
  if (!dlopen(path, RTLD_NOW | RTLD_GLOBAL)) die(dlerror());
  snprintf(path, sizeof path, "%s/libXext.so.6", self);
  if (!dlopen(path, RTLD_NOW | RTLD_GLOBAL)) die(dlerror());
  void* gl = dlopen(libgl, RTLD_NOW | RTLD_GLOBAL);
  if (!gl) die(dlerror());

  // ---- an off-screen GLX context on the stand-in display
  static Display dpy;
  static Screen scr;
  static Depth depth;
  static ScreenFormat fmt = {0, 24, 32, 32};
  Visual* vis = ((Visual * (*)(void)) dlsym(RTLD_DEFAULT, "fakex_visual"))();
  depth.depth = 24; depth.nvisuals = 1; depth.visuals = vis;
  scr.display = &dpy; scr.root = 1; scr.width = 1024; scr.height = 768; scr.ndepths = 1; scr.depths = &depth; scr.root_depth = 24;
  scr.root_visual = vis; scr.cmap = 0x31;
  dpy.bitmap_unit = 32; dpy.bitmap_pad = 32; dpy.nformats = 1; dpy.pixmap_format = &fmt; dpy.nscreens = 1; dpy.screens = &scr;
  dpy.display_name = ":fake"; dpy.vendor = "fake";
  void* (*gpa)(const char*) = dlsym(gl, "glXGetProcAddress");
  GLXFBConfig* (*choose)(Display*, int, const int*, int*) = dlsym(gl, "glXChooseFBConfig");
  GLXPbuffer (*mkpb)(Display*, GLXFBConfig, const int*) = dlsym(gl, "glXCreatePbuffer");
  GLXContext (*mkctx)(Display*, GLXFBConfig, int, GLXContext, int) = dlsym(gl, "glXCreateNewContext");
  int (*mc)(Display*, GLXPbuffer, GLXPbuffer, GLXContext) = dlsym(gl, "glXMakeContextCurrent");
  if (!gpa || !choose || !mkpb || !mkctx || !mc) die("GLX entry points missing");
  int attribs[] = {0x8010 /*DRAWABLE_TYPE*/, 0x4 /*PBUFFER*/, 0x8011 /*RENDER_TYPE*/, 1, 8, 8, 9, 8, 10, 8, 12, 24, 0}, ncfg = 0;
  GLXFBConfig* cfg = choose(&dpy, 0, attribs, &ncfg);
  if (!cfg || !ncfg) die("no GLX framebuffer config");
  int pba[] = {0x8041, 64, 0x8040, 64, 0};
  GLXPbuffer pb = mkpb(&dpy, cfg[0], pba);
  GLXContext ctx = mkctx(&dpy, cfg[0], 0x8014 /*RGBA_TYPE*/, 0, 1);
  if (!ctx || !mc(&dpy, pb, pb, ctx)) die("cannot create / bind the GL context");
#define X(ret, name, args) gl##name = (ret(*) args)gpa("gl" #name); if (!gl##name) die("missing gl" #name);
  GL_FUNCS(X)
#undef X
  fprintf(stderr, "GL_VERSION %s | GL_RENDERER %s | GLSL %s\n", glGetString(0x1F02), glGetString(0x1F01), glGetString(0x8B8C));

  // ---- the job
  FILE* f = fopen(argv[2], "rb");
  if (!f) die("cannot open the job file");
  char magic[8];
  if (fread(magic, 1, 8, f) == 8 && !memcmp(magic, "VCTGLMIP", 8)) { fclose(f); return mip_pass(argv); }
  rewind(f);
  struct Job J;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "VCTGLJOB", 8) || fread(&J, sizeof J, 1, f) != 1) die("bad job file");
#define READ(ptr, bytes) do { const size_t nb_ = (bytes); (ptr) = malloc(nb_ + 1); if (fread((ptr), 1, nb_, f) != nb_) die("short job file"); } while (0)
  float* lights; READ(lights, (size_t)J.n_lights * 28);
  unsigned char* mats; READ(mats, (size_t)J.n_mats * 128);
  unsigned char* draws; READ(draws, (size_t)J.n_draws * 80);
  unsigned char* verts; READ(verts, (size_t)J.n_verts * 32);
  unsigned char* indices; READ(indices, (size_t)J.n_indices * 4);

  // ---- buffers: one VAO over all vertices / indices (load_model, renderer.cpp:559-600), one UBO per material, the camera UBO
  GLuint vao, vbo, ebo, cam_ubo;
  glGenVertexArrays(1, &vao);
  glBindVertexArray(vao);
  glGenBuffers(1, &vbo); glBindBuffer(0x8892 /*ARRAY_BUFFER*/, vbo); glBufferData(0x8892, (GLsizeiptr)J.n_verts * 32, verts, 0x88E4);
  glGenBuffers(1, &ebo); glBindBuffer(0x8893 /*ELEMENT_ARRAY_BUFFER*/, ebo); glBufferData(0x8893, (GLsizeiptr)J.n_indices * 4, indices, 0x88E4);
  glEnableVertexAttribArray(0); glVertexAttribPointer(0, 3, 0x1406, 0, 32, (const void*)0);
  glEnableVertexAttribArray(1); glVertexAttribPointer(1, 3, 0x1406, 0, 32, (const void*)12);
  glEnableVertexAttribArray(2); glVertexAttribPointer(2, 2, 0x1406, 0, 32, (const void*)24);
  GLuint* mat_ubo = malloc(sizeof(GLuint) * (J.n_mats ? J.n_mats : 1));
  glGenBuffers((GLsizei)J.n_mats, mat_ubo);
  for (uint32_t i = 0; i < J.n_mats; i++) { glBindBuffer(0x8A11 /*UNIFORM_BUFFER*/, mat_ubo[i]); glBufferData(0x8A11, 128, mats + 128 * (size_t)i, 0x88E4); }
  float cam[32];
  memcpy(cam, J.proj, 64); memcpy(cam + 16, J.view, 64);   // camera_data_t: projection, view (renderer.h:81-85)
  glGenBuffers(1, &cam_ubo); glBindBuffer(0x8A11, cam_ubo); glBufferData(0x8A11, 128, cam, 0x88E8);

  if (J.levels == 0) {   // voxelization pass (see the header): fragments of every triangle, no textures in the job
    fclose(f);
    return voxelize_pass(&J, argv, lights, draws, mat_ubo, cam_ubo);
  }

  // ---- create_tex_3d (texture_3d.cpp:3-25): RGBA8, `levels` levels, CLAMP_TO_BORDER, LINEAR_MIPMAP_LINEAR (the MAG_FILTER call with
  // that enum is an error in GL and leaves GL_LINEAR, as in the reference); texels from the job file
  GLuint tex[6];
  glGenTextures(6, tex);
  glPixelStorei(0x0CF5 /*UNPACK_ALIGNMENT*/, 1);
  size_t n0 = (size_t)J.R * J.R * J.R;
  unsigned char* level0; READ(level0, n0 * 4);
  for (int d = 0; d < 6; d++) {
    glBindTexture(0x806F /*TEXTURE_3D*/, tex[d]);
    glTexParameteri(0x806F, 0x2802 /*WRAP_S*/, 0x812D /*CLAMP_TO_BORDER*/);
    glTexParameteri(0x806F, 0x2803 /*WRAP_T*/, 0x812D);
    glTexParameteri(0x806F, 0x8072 /*WRAP_R*/, 0x812D);
    glTexParameteri(0x806F, 0x2801 /*MIN_FILTER*/, 0x2703 /*LINEAR_MIPMAP_LINEAR*/);
    glTexParameteri(0x806F, 0x2800 /*MAG_FILTER*/, 0x2703);
    (void)glGetError();   // INVALID_ENUM from the line above
    glTexStorage3D(0x806F, (GLsizei)J.levels, 0x8058 /*RGBA8*/, J.R, J.R, J.R);
    glTexSubImage3D(0x806F, 0, 0, 0, 0, J.R, J.R, J.R, 0x1908 /*RGBA*/, 0x1401 /*UNSIGNED_BYTE*/, level0);
    for (uint32_t l = 1; l < J.levels; l++) {
      const uint32_t N = J.R >> l;
      unsigned char* texels; READ(texels, (size_t)N * N * N * 4);
      glTexSubImage3D(0x806F, (GLint)l, 0, 0, 0, N, N, N, 0x1908, 0x1401, texels);
      free(texels);
    }
  }
  fclose(f);
  if (glGetError()) die("GL error while creating the voxel textures");

  // ---- load_shader (renderer.cpp:605-641): the reference's text, unmodified
  GLuint prog = glCreateProgram();
  glAttachShader(prog, compile(0x8B31, slurp(argv[1], "voxel_cone_tracing.vert"), "voxel_cone_tracing.vert"));
  glAttachShader(prog, compile(0x8B30, slurp(argv[1], "voxel_cone_tracing.frag"), "voxel_cone_tracing.frag"));
  glLinkProgram(prog);
  GLint ok = 0;
  glGetProgramiv(prog, 0x8B82, &ok);
  if (!ok) { char log[8192] = ""; glGetProgramInfoLog(prog, sizeof log, 0, log); fprintf(stderr, "%s\n", log); die("link failed"); }
  const GLuint material_location = glGetUniformBlockIndex(prog, "material"), camera_location = glGetUniformBlockIndex(prog, "camera");
  fprintf(stderr, "uniform block indices: camera %u, material %u (the reference binds its UBOs at these numbers; the shaders say binding 0 / 1)\n",
          camera_location, material_location);

  FILE* out = fopen(argv[3], "wb");
  if (!out) die("cannot open the output file");
  for (int pass = 0; pass < 2; pass++) {   // 0: RGBA8 colour buffer (the window of the reference), 1: RGBA32F (the shader's values unrounded)
    GLuint fbo, rb[2];
    glGenFramebuffers(1, &fbo); glBindFramebuffer(0x8D40, fbo);
    glGenRenderbuffers(2, rb);
    glBindRenderbuffer(0x8D41, rb[0]); glRenderbufferStorage(0x8D41, pass ? 0x8814 /*RGBA32F*/ : 0x8058 /*RGBA8*/, J.W, J.H);
    glFramebufferRenderbuffer(0x8D40, 0x8CE0 /*COLOR_ATTACHMENT0*/, 0x8D41, rb[0]);
    glBindRenderbuffer(0x8D41, rb[1]); glRenderbufferStorage(0x8D41, 0x81A6 /*DEPTH_COMPONENT24*/, J.W, J.H);
    glFramebufferRenderbuffer(0x8D40, 0x8D00 /*DEPTH_ATTACHMENT*/, 0x8D41, rb[1]);
    if (glCheckFramebufferStatus(0x8D40) != 0x8CD5) die("framebuffer incomplete");
    if (pass) { glClampColor(0x891B /*CLAMP_FRAGMENT_COLOR*/, 0); glClampColor(0x891C /*CLAMP_READ_COLOR*/, 0); }

    // ---- Renderer::render / visualize (renderer.cpp:355-400)
    glClearColor(0.15f, 0.25f, 0.25f, 1.0f);
    glViewport(0, 0, J.W, J.H);
    glEnable(0x0B71 /*DEPTH_TEST*/);
    glClear(0x4000 | 0x100);
    glUseProgram(prog);
    glUniform1f(glGetUniformLocation(prog, "cube_size"), J.cube_size);
    glUniform1i(glGetUniformLocation(prog, "cube_res"), (GLint)J.R);
    glUniform1i(glGetUniformLocation(prog, "enable_diffuse"), J.diffuse);
    glUniform1i(glGetUniformLocation(prog, "enable_specular"), J.specular);
    glUniform1i(glGetUniformLocation(prog, "enable_shadow"), J.shadow);
    glUniform1i(glGetUniformLocation(prog, "enable_direct"), J.direct);
    glUniform1i(glGetUniformLocation(prog, "view_voxel_dir"), J.view_voxel_dir);
    glUniform1f(glGetUniformLocation(prog, "view_voxel_lod"), J.view_voxel_lod);
    for (int i = 0; i < 6; i++) { glActiveTexture(0x84C0 + 2 + i); glBindTexture(0x806F, tex[i]); }   // activate_tex_3d(.., i + 2)
    glBindBufferBase(0x8A11, camera_location, cam_ubo);                                                // upload_camera
    glUniform3fv(glGetUniformLocation(prog, "camera_position"), 1, J.view + 12);                       // glm::column(view, 3)
    for (uint32_t i = 0; i < J.n_lights; i++) {                                                        // upload_lights
      char name[64];
      snprintf(name, sizeof name, "point_lights[%u].position", i); glUniform3fv(glGetUniformLocation(prog, name), 1, lights + 7 * i);
      snprintf(name, sizeof name, "point_lights[%u].color", i); glUniform3fv(glGetUniformLocation(prog, name), 1, lights + 7 * i + 3);
      snprintf(name, sizeof name, "point_lights[%u].intensity", i); glUniform1f(glGetUniformLocation(prog, name), lights[7 * i + 6]);
    }
    glUniform1i(glGetUniformLocation(prog, "point_light_count"), (GLint)J.n_lights);
    glEnable(0x0BE2 /*BLEND*/);
    glBlendFunc(0x0302 /*SRC_ALPHA*/, 0x0303 /*ONE_MINUS_SRC_ALPHA*/);
    const GLint model_location = glGetUniformLocation(prog, "model");
    for (uint32_t i = 0; i < J.n_draws; i++) {                                                         // draw_models
      uint32_t d[4];
      memcpy(d, draws + 80 * (size_t)i, 16);
      glUniformMatrix4fv(model_location, 1, 0, (const float*)(draws + 80 * (size_t)i + 16));
      if (d[3] >= J.n_mats) die("draw refers to an unknown material");
      glBindBufferBase(0x8A11, material_location, mat_ubo[d[3]]);                                      // bind_material
      // one VAO holds every model here: the model's vertex base is added by the draw call instead of by a VAO of its own
      glDrawElementsBaseVertex(0x0004 /*TRIANGLES*/, (GLsizei)d[1], 0x1405 /*UNSIGNED_INT*/, (const void*)(sizeof(unsigned) * (size_t)d[0]), (GLint)d[2]);
    }
    glFinish();
    if (glGetError()) die("GL error while drawing");
    const size_t px = (size_t)J.W * J.H;
    void* buf = malloc(px * (pass ? 16 : 4));
    glPixelStorei(0x0D05 /*PACK_ALIGNMENT*/, 1);
    glReadPixels(0, 0, J.W, J.H, 0x1908, pass ? 0x1406 /*FLOAT*/ : 0x1401, buf);
    if (glGetError()) die("GL error in glReadPixels");
    fwrite(buf, pass ? 16 : 4, px, out);
    free(buf);
  }
  fclose(out);
  return 0;
}
