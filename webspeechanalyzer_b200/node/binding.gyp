{
  "targets": [{
    "target_name": "fa_b200",
    "sources": ["fa_napi.c"],
    "defines": ["FA_USE_SYSTEM_NODE_API"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/..", "-lfa_b200", "-Wl,-rpath,<(module_root_dir)/.."]
  }]
}
