P=mulit_view_object_detection_b200
cp $P/libmvfusion.so /tmp/lib_prod.so
for f in libmvfusion libmvfusion_prof libmvfusion_dbg libmvfusion_head; do
  echo "== $f"; cp $P/$f.so /tmp/cur.so; cp /tmp/cur.so $P/libmvfusion.so 2>/dev/null || true
  if [ $f = libmvfusion ]; then cp /tmp/lib_prod.so $P/libmvfusion.so; fi
  timeout 100 python tools/k1t_debug.py timing 2>&1 | grep "tensor_cores=True"
done
cp /tmp/lib_prod.so $P/libmvfusion.so
