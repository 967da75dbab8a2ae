{
  "targets": [{
    "target_name": "fspt_napi",
    "sources": ["fspt_napi.cc"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../lib", "-lfspt_b200", "-Wl,-rpath,<(module_root_dir)/../lib"],
    "cflags_cc": ["-std=c++17"]
  }]
}
