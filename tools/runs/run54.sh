for c in 2 4 8; do echo "SDOD_GN_CS=$c"; SDOD_GN_CS=$c timeout 200 python tools/gn_c1.py 2>&1 | tail -1 | cut -c95-260; done
