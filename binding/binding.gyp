{
  # node-gyp rebuild   (run inside binding/, with libjslp_b200.so built by `python -m jslpsolver_b200.build`)
  "targets": [
    {
      "target_name": "jslp_b200",
      "sources": ["jslp_addon.cc"],
      "include_dirs": ["../include"],
      "cflags_cc": ["-std=c++17", "-O2", "-Wall"],
      "cflags_cc!": ["-fno-exceptions"],
      "defines": ["NAPI_VERSION=6"],
      "libraries": ["-L<(module_root_dir)/../jslpsolver_b200", "-ljslp_b200",
                    "-Wl,-rpath,<(module_root_dir)/../jslpsolver_b200"]
    }
  ]
}
