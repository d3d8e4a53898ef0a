timeout 300 python tools/defer_diag.py lego_render 96 2>&1 | tail -20
timeout 300 python tools/defer_diag.py lego_render 200 2>&1 | tail -16
